// topolow_b200/csrc/post.cu
//
// Post-processing of one fit on the device:
//   topolow_est_distances   as.matrix(stats::dist(positions))            R/core.R:474
//   topolow_holdout_errors  sum |truth - est| over the held-out cells, the OutSampleError
//                           reduction of R/error_metrics.R:95-114 as pooled by
//                           R/adaptive_sampling.R:2642-2647
// FP64 throughout (R computes both in double).
#include <vector>

#include "../../include/topolow_b200.h"
#include "common.cuh"

namespace tl {
namespace {

// pos: [n][dim] row-major.  One CTA per 32x32 output tile, both point tiles staged in smem.
__global__ void __launch_bounds__(256) dist_kernel(const double* __restrict__ pos, int64_t n, int dim,
                                                   double* __restrict__ out) {
  extern __shared__ double sm[];
  double* sa = sm;                 // [32][dim]
  double* sb = sm + 32 * dim;      // [32][dim]
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int x = threadIdx.x; x < 32 * dim; x += blockDim.x) {
    const int64_t ra = r0 + x / dim, rb = c0 + x / dim;
    sa[x] = ra < n ? pos[ra * dim + x % dim] : 0.0;
    sb[x] = rb < n ? pos[rb * dim + x % dim] : 0.0;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < 1024; x += blockDim.x) {
    const int i = x / 32, j = x % 32;
    const int64_t r = r0 + i, c = c0 + j;
    if (r >= n || c >= n) continue;
    double ss = 0.0;
    for (int d = 0; d < dim; ++d) {
      const double df = __dsub_rn(sa[i * dim + d], sb[j * dim + d]);
      ss = __dadd_rn(ss, __dmul_rn(df, df));
    }
    out[r * n + c] = __dsqrt_rn(ss);
  }
}

__global__ void __launch_bounds__(256) holdout_kernel(const double* __restrict__ pos, int dim, int64_t n_cells,
                                                      const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                                                      const double* __restrict__ truth, double* __restrict__ part_sum,
                                                      unsigned long long* __restrict__ part_cnt) {
  double s = 0.0;
  unsigned long long c = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_cells; e += (int64_t)gridDim.x * blockDim.x) {
    const double t = truth[e];
    if (isnan(t)) continue;  // NA truth cells are dropped (R/adaptive_sampling.R:2642)
    const double* a = pos + (int64_t)ci[e] * dim;
    const double* b = pos + (int64_t)cj[e] * dim;
    double ss = 0.0;
    for (int d = 0; d < dim; ++d) {
      const double df = __dsub_rn(a[d], b[d]);
      ss = __dadd_rn(ss, __dmul_rn(df, df));
    }
    s += fabs(t - __dsqrt_rn(ss));
    c += 1;
  }
  __shared__ double ss_[256];
  __shared__ unsigned long long cc_[256];
  ss_[threadIdx.x] = s; cc_[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { ss_[threadIdx.x] += ss_[threadIdx.x + o]; cc_[threadIdx.x] += cc_[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_sum[blockIdx.x] = ss_[0]; part_cnt[blockIdx.x] = cc_[0]; }
}

// The same reduction on positions that are still on the device in slot order (`real` = float or double;
// a float converts to double exactly, so the terms equal those of holdout_kernel on the downloaded
// positions, and the grid and the summation order are the same).
template <class real>
__global__ void __launch_bounds__(256) holdout_slots_kernel(const real* __restrict__ pos, int dim, int64_t n_cells,
                                                            const int32_t* __restrict__ si, const int32_t* __restrict__ sj,
                                                            const double* __restrict__ truth, double* __restrict__ part_sum,
                                                            unsigned long long* __restrict__ part_cnt) {
  double s = 0.0;
  unsigned long long c = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_cells; e += (int64_t)gridDim.x * blockDim.x) {
    const double t = truth[e];
    if (isnan(t)) continue;
    const real* a = pos + (int64_t)si[e] * dim;
    const real* b = pos + (int64_t)sj[e] * dim;
    double ss = 0.0;
    for (int d = 0; d < dim; ++d) {
      const double df = __dsub_rn((double)a[d], (double)b[d]);
      ss = __dadd_rn(ss, __dmul_rn(df, df));
    }
    s += fabs(t - __dsqrt_rn(ss));
    c += 1;
  }
  __shared__ double ss_[256];
  __shared__ unsigned long long cc_[256];
  ss_[threadIdx.x] = s; cc_[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { ss_[threadIdx.x] += ss_[threadIdx.x + o]; cc_[threadIdx.x] += cc_[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_sum[blockIdx.x] = ss_[0]; part_cnt[blockIdx.x] = cc_[0]; }
}

std::vector<double> to_row_major(const double* cm, int64_t n, int dim) {
  std::vector<double> rm((size_t)n * dim);
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) rm[(size_t)i * dim + d] = cm[(size_t)d * n + i];
  return rm;
}

}  // namespace
}  // namespace tl

extern "C" int topolow_est_distances(const double* positions, int64_t n, int32_t ndim, double* est, int32_t device) {
  using namespace tl;
  if (!positions || !est || n < 1 || ndim < 1 || ndim > 1024) return TOPOLOW_ERR_BAD_ARG;
  try {
    TL_CUDA(cudaSetDevice(device));
    const std::vector<double> rm = to_row_major(positions, n, ndim);
    DeviceBuf<double> d_pos(rm.size()), d_out((size_t)n * n);
    TL_CUDA(cudaMemcpy(d_pos, rm.data(), rm.size() * sizeof(double), cudaMemcpyHostToDevice));
    const unsigned nb = (unsigned)((n + 31) / 32);
    dist_kernel<<<dim3(nb, nb), 256, 2 * 32 * ndim * sizeof(double)>>>(d_pos, n, ndim, d_out);
    TL_CUDA(cudaGetLastError());
    TL_CUDA(cudaMemcpy(est, d_out, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
    return TOPOLOW_OK;
  } catch (const CudaError&) {
    cudaGetLastError();
    return TOPOLOW_ERR_CUDA;
  }
}

namespace tl {
// Hold-out residuals of a finished fit from its device-resident best positions ([slot][dim], FP32 or FP64).
void holdout_resident(const void* best, bool is_f64, int dim, int64_t n_cells, const int32_t* d_slot_i,
                      const int32_t* d_slot_j, const double* d_truth, cudaStream_t stream, double* sum_abs,
                      int64_t* count) {
  const int blocks = 296;
  AsyncBuf<double> d_ps(blocks, stream);
  AsyncBuf<unsigned long long> d_pc(blocks, stream);
  if (is_f64) holdout_slots_kernel<double><<<blocks, 256, 0, stream>>>((const double*)best, dim, n_cells, d_slot_i, d_slot_j, d_truth, d_ps, d_pc);
  else holdout_slots_kernel<float><<<blocks, 256, 0, stream>>>((const float*)best, dim, n_cells, d_slot_i, d_slot_j, d_truth, d_ps, d_pc);
  TL_CUDA(cudaGetLastError());
  std::vector<double> ps(blocks);
  std::vector<unsigned long long> pc(blocks);
  TL_CUDA(cudaMemcpyAsync(ps.data(), d_ps, blocks * sizeof(double), cudaMemcpyDeviceToHost, stream));
  TL_CUDA(cudaMemcpyAsync(pc.data(), d_pc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  TL_CUDA(cudaStreamSynchronize(stream));
  double s = 0.0; unsigned long long c = 0;
  for (int b = 0; b < blocks; ++b) { s += ps[b]; c += pc[b]; }
  *sum_abs = s; *count = (int64_t)c;
}
}  // namespace tl

extern "C" int topolow_holdout_errors(const double* positions, int64_t n, int32_t ndim, int64_t n_cells,
                                      const int32_t* cell_i, const int32_t* cell_j, const double* truth,
                                      double* sum_abs_out, int64_t* count_out, int32_t device) {
  using namespace tl;
  if (!positions || n < 1 || ndim < 1 || n_cells < 0 || !sum_abs_out || !count_out) return TOPOLOW_ERR_BAD_ARG;
  for (int64_t e = 0; e < n_cells; ++e)
    if (cell_i[e] < 0 || cell_j[e] < 0 || cell_i[e] >= n || cell_j[e] >= n) return TOPOLOW_ERR_BAD_ARG;
  try {
    TL_CUDA(cudaSetDevice(device));
    const std::vector<double> rm = to_row_major(positions, n, ndim);
    const int blocks = 296;
    // (pool scratch: a CV grid calls this once per fit)
    cudaStream_t ps_ = cudaStreamPerThread;
    AsyncBuf<double> d_pos(rm.size(), ps_), d_truth(n_cells, ps_), d_ps(blocks, ps_);
    AsyncBuf<int32_t> d_ci(n_cells, ps_), d_cj(n_cells, ps_);
    AsyncBuf<unsigned long long> d_pc(blocks, ps_);
    TL_CUDA(cudaStreamSynchronize(ps_));
    TL_CUDA(cudaMemcpy(d_pos, rm.data(), rm.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (n_cells > 0) {
      TL_CUDA(cudaMemcpy(d_truth, truth, n_cells * sizeof(double), cudaMemcpyHostToDevice));
      TL_CUDA(cudaMemcpy(d_ci, cell_i, n_cells * sizeof(int32_t), cudaMemcpyHostToDevice));
      TL_CUDA(cudaMemcpy(d_cj, cell_j, n_cells * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    holdout_kernel<<<blocks, 256>>>(d_pos, ndim, n_cells, d_ci, d_cj, d_truth, d_ps, d_pc);
    TL_CUDA(cudaGetLastError());
    std::vector<double> ps(blocks);
    std::vector<unsigned long long> pc(blocks);
    TL_CUDA(cudaMemcpy(ps.data(), d_ps, blocks * sizeof(double), cudaMemcpyDeviceToHost));
    TL_CUDA(cudaMemcpy(pc.data(), d_pc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double s = 0.0; unsigned long long c = 0;
    for (int b = 0; b < blocks; ++b) { s += ps[b]; c += pc[b]; }
    *sum_abs_out = s; *count_out = (int64_t)c;
    return TOPOLOW_OK;
  } catch (const CudaError&) {
    cudaGetLastError();
    return TOPOLOW_ERR_CUDA;
  }
}
