// topolow_b200/csrc/microbench.cu
//
// Pipe-rate probes used as roofline denominators (the FP32 FMA peak is
// not in MEASURED_PEAKS.json and has to be measured on the box).  Each probe runs a
// register-resident dependent-chain kernel sized to saturate every SM and is timed with
// CUDA events on its own stream.
#include "../../include/topolow_b200.h"
#include "common.cuh"

namespace tl {
namespace {

constexpr int kChains = 8;      // independent chains per thread (covers the 4-cycle latency)
constexpr int kInner = 4096;    // loop trips

__global__ void __launch_bounds__(256) ffma_kernel(float* out, float a, float b) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 1e-3f + c;
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = fmaf(x[c], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float* out, float a, float b) {
  unsigned long long x[kChains];
  unsigned long long pa, pb;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    float v = threadIdx.x * 1e-3f + c;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(x[c]) : "f"(v));
  }
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(pa), "l"(pb));
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[c]));
    s += lo + hi;
  }
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0.;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) shfl_kernel(float* out) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x + c;
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = __shfl_sync(0xffffffffu, x[c], (threadIdx.x + 1) & 31);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) rsq_kernel(float* out) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x + c + 1.5f;
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = rsqrtf(x[c]) + 1.0f;  // MUFU.RSQ + FADD
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456f) out[0] = s;
}

// FFMA2 and SHFL streams in the same warp (independent chains): does the shuffle unit run beside the
// FMA pipe, or do the two share the sub-partition's dispatch slot?
__global__ void __launch_bounds__(256) mix_kernel(float* out, float a, float b, int n_ffma2, int n_shfl) {
  float2 x[kChains];
  float y[kChains];
  const float2 pa = make_float2(a, a), pb = make_float2(b, b);
#pragma unroll
  for (int c = 0; c < kChains; ++c) { x[c] = make_float2(threadIdx.x * 1e-3f + c, c); y[c] = threadIdx.x + c; }
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
      if (c < n_ffma2) x[c] = __ffma2_rn(x[c], pa, pb);
      if (c < n_shfl) y[c] = __shfl_sync(0xffffffffu, y[c], (threadIdx.x + 1) & 31);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c].x + x[c].y + y[c];
  if (s == 123.456f) out[0] = s;
}

// Legacy warp-level tensor instruction (mma.sync m16n8k8, TF32 inputs, FP32 accumulate): kChains independent accumulator
// tiles per warp.  n_mma / n_fma / n_mufu per trip let the probe mix it with FP32 and special-function work (do the
// tensor pipe and the FMA / XU pipes overlap for a warp that issues in order?).
__global__ void __launch_bounds__(256) mma_tf32_kernel(float* out, int n_mma, int n_fma, int n_mufu) {
  float c[kChains][4];
  float f[kChains];
  unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), a1 = a0 + 64, a2 = a0 + 128, a3 = a0 + 192;
  unsigned b0 = __float_as_uint(0.5f), b1 = __float_as_uint(0.25f);
#pragma unroll
  for (int k = 0; k < kChains; ++k) { c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.f; f[k] = 1.0f + k; }
  for (int i = 0; i < kInner; ++i) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
      if (k < n_mma)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < kChains; ++k)
        if (j * kChains + k < n_fma) f[k] = fmaf(f[k], 1.0001f, 0.5f);
#pragma unroll
    for (int k = 0; k < kChains; ++k)
      if (k < n_mufu) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3] + f[k];
  if (s == 123.456f) out[0] = s;
}

__global__ void copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) out[i] = in[i];
}

}  // namespace
}  // namespace tl

extern "C" int topolow_microbench(int32_t which, int32_t device, double* value_out) {
  using namespace tl;
  try {
    TL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop; TL_CUDA(cudaGetDeviceProperties(&prop, device));
    cudaStream_t s; TL_CUDA(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; TL_CUDA(cudaEventCreate(&e0)); TL_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    float* d_out = nullptr; TL_CUDA(cudaMalloc(&d_out, 64));
    float4 *d_a = nullptr, *d_b = nullptr;
    const size_t copy_bytes = 1ull << 30;
    if (which == 5) { TL_CUDA(cudaMalloc(&d_a, copy_bytes)); TL_CUDA(cudaMalloc(&d_b, copy_bytes));
                      TL_CUDA(cudaMemset(d_a, 1, copy_bytes)); }
    float best_ms = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      TL_CUDA(cudaEventRecord(e0, s));
      switch (which) {
        case 0: ffma_kernel<<<blocks, threads, 0, s>>>(d_out, 1.0001f, 0.5f); break;
        case 1: ffma2_kernel<<<blocks, threads, 0, s>>>(d_out, 1.0001f, 0.5f); break;
        case 2: dfma_kernel<<<blocks, threads, 0, s>>>((double*)d_out, 1.0001, 0.5); break;
        case 3: shfl_kernel<<<blocks, threads, 0, s>>>(d_out); break;
        case 4: rsq_kernel<<<blocks, threads, 0, s>>>(d_out); break;
        case 5: copy_kernel<<<prop.multiProcessorCount * 16, 512, 0, s>>>(d_a, d_b, copy_bytes / 16); break;
        case 6: mix_kernel<<<blocks, threads, 0, s>>>(d_out, 1.0001f, 0.5f, 8, 0); break;   // 8 FFMA2 per trip
        case 7: mix_kernel<<<blocks, threads, 0, s>>>(d_out, 1.0001f, 0.5f, 0, 4); break;   // 4 SHFL per trip
        case 8: mix_kernel<<<blocks, threads, 0, s>>>(d_out, 1.0001f, 0.5f, 8, 4); break;   // both
        case 9: mma_tf32_kernel<<<blocks, threads, 0, s>>>(d_out, 8, 0, 0); break;                                  // tensor only, full occupancy
        case 10: mma_tf32_kernel<<<prop.multiProcessorCount * 2, threads, 0, s>>>(d_out, 8, 0, 0); break;         // 4 warps per scheduler
        case 11: mma_tf32_kernel<<<prop.multiProcessorCount * 2, threads, 0, s>>>(d_out, 5, 24, 6); break;        // the mix of one trip
        case 12: mma_tf32_kernel<<<prop.multiProcessorCount * 2, threads, 0, s>>>(d_out, 0, 24, 6); break;        // its FP32 / MUFU part
        case 13: mma_tf32_kernel<<<prop.multiProcessorCount * 2, threads, 0, s>>>(d_out, 5, 0, 0); break;         // its tensor part
        default: return TOPOLOW_ERR_BAD_ARG;
      }
      TL_CUDA(cudaGetLastError());
      TL_CUDA(cudaEventRecord(e1, s));
      TL_CUDA(cudaEventSynchronize(e1));
      float ms; TL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    const double sec = best_ms * 1e-3;
    const double thread_ops = (double)blocks * threads * kInner * kChains;
    double v = 0;
    if (which == 0) v = thread_ops * 2.0 / sec;            // flop/s
    else if (which == 1) v = thread_ops * 4.0 / sec;       // flop/s (2 lanes per instruction)
    else if (which == 2) v = thread_ops * 2.0 / sec;       // flop/s
    else if (which == 3) v = thread_ops / 32.0 / sec;      // warp instructions / s
    else if (which == 4) v = thread_ops / 32.0 / sec;      // warp MUFU instructions / s
    else if (which == 5) v = 2.0 * copy_bytes / sec;       // bytes/s
    else if (which == 9) v = thread_ops / 32.0 * (2.0 * 16 * 8 * 8) / sec;                                        // flop/s
    else if (which == 10) v = (double)prop.multiProcessorCount * 2 * threads * kInner * kChains / 32.0 * (2.0 * 16 * 8 * 8) / sec;
    else v = sec * 1e3;                                    // 6..8, 11..13: milliseconds of the loop
    *value_out = v;
    cudaFree(d_out); if (d_a) cudaFree(d_a); if (d_b) cudaFree(d_b);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
    return TOPOLOW_OK;
  } catch (const CudaError&) {
    return TOPOLOW_ERR_CUDA;
  }
}
