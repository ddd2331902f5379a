// tools/ptxas_convergence_probe.cu - when does ptxas guard shuffles with BRA.DIV?  
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -c tools/ptxas_convergence_probe.cu -o /tmp/p.o && cuobjdump -sass /tmp/p.o | grep -E "Function|BRA.DIV"
// Observed (CUDA 12.9): only k7 - shuffles under a branch on a per-lane loaded value - gets the guard; uniform-address
// loads, vote results, shuffle results and structured divergence that has reconverged do not.
#include <cuda_runtime.h>
__global__ void k0(float* out, const int* cnt) {   // trip count loaded from memory
  float x = threadIdx.x;
  int n = cnt[0];
  for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
  out[threadIdx.x] = x;
}
__global__ void k1(float* out, const int* cnt) {   // loaded per lane, made uniform through a vote
  float x = threadIdx.x;
  int n = cnt[threadIdx.x];
  for (int i = 0; __any_sync(0xffffffffu, i < n); ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
  out[threadIdx.x] = x;
}
__global__ void k2(float* out, volatile int* flag) {   // spin wait then shuffle
  float x = threadIdx.x;
  while (*flag < 3) { }
  __syncwarp();
  for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
  out[threadIdx.x] = x;
}
__global__ void k3(float* out, volatile int* flag) {   // spin wait with vote exit
  float x = threadIdx.x;
  while (!__all_sync(0xffffffffu, *flag >= 3)) { }
  for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
  out[threadIdx.x] = x;
}
__global__ void k4(float* out, const int* cnt) {   // per-lane data-dependent if, then shuffles after reconvergence
  float x = threadIdx.x;
  if (cnt[threadIdx.x] > 3) x = sqrtf(x) + out[5];
  for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
  out[threadIdx.x] = x;
}
__global__ void k5(float* out, const int* cnt) {   // shuffle inside if with shuffled (uniform at run time) condition
  float x = threadIdx.x;
  int c = __shfl_sync(0xffffffffu, cnt[threadIdx.x], 0);
  if (c > 3) { for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f; }
  out[threadIdx.x] = x;
}
__global__ void k6(float* out, const int* cnt) {   // same with vote-made condition
  float x = threadIdx.x;
  if (__any_sync(0xffffffffu, cnt[threadIdx.x] > 3)) { for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f; }
  out[threadIdx.x] = x;
}
__global__ void k7(float* out, const int* cnt) {   // shuffles under a per-lane loaded condition
  float x = threadIdx.x;
  if (cnt[threadIdx.x] > 3) { for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f; }
  out[threadIdx.x] = x;
}
__global__ void k8(float* out, const int* cnt) {   // loop with break on loaded value
  float x = threadIdx.x;
  __shared__ int st;
  if (threadIdx.x == 0) st = cnt[0];
  __syncthreads();
  for (int it = 0; it < 100; ++it) {
    if (st) break;
    for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
    if (threadIdx.x == 0) st = x > 100.f;
    __syncthreads();
  }
  out[threadIdx.x] = x;
}
__global__ void k9(float* out, const int* cnt) {   // same, vote-made break
  float x = threadIdx.x;
  __shared__ int st;
  if (threadIdx.x == 0) st = cnt[0];
  __syncthreads();
  for (int it = 0; it < 100; ++it) {
    if (__any_sync(0xffffffffu, st)) break;
    for (int i = 0; i < 8; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.5f;
    if (threadIdx.x == 0) st = x > 100.f;
    __syncthreads();
  }
  out[threadIdx.x] = x;
}
