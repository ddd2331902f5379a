// topolow_b200/csrc/graph.cu
//
// Connected components of the measurement graph, for many candidate subsamples at once:
//   topolow_components   what check_matrix_connectivity asks igraph for (R/utils.R:199-253: adjacency = !is.na off
//                        the diagonal, R/diagnostics.R:434-472; igraph::components()$no) on the sub-matrix selected by
//                        each of n_masks point masks - the attempts of subsample_dissimilarity_matrix
//                        (R/utils.R:371-440) - from ONE uploaded edge list.
// Union-find on the edge list: every (mask, edge) thread finds the roots of its endpoints and hooks the larger root
// under the smaller one with atomicMin; a pointer-jumping pass flattens the trees; repeated until a pass changes
// nothing (a hook that loses the race is retried in the next pass: its edge still joins two different roots).
#include <vector>

#include "../../include/topolow_b200.h"
#include "common.cuh"

namespace tl {
namespace {

__device__ __forceinline__ int find_root(const int* __restrict__ parent, int v) {
  int p = __ldcg(parent + v);
  while (p != v) { v = p; p = __ldcg(parent + v); }
  return v;
}

__global__ void cc_init_kernel(int* parent, long long total, int n) {
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x)
    parent[x] = (int)(x % n);
}

__global__ void cc_hook_kernel(int* parent, int n, long long E, const int32_t* __restrict__ ei, const int32_t* __restrict__ ej,
                               int n_masks, const unsigned char* __restrict__ masks, int* changed) {
  const long long total = (long long)n_masks * E;
  bool any = false;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(x / E);
    const long long e = x % E;
    const int a = ei[e], b = ej[e];
    if (masks && !(masks[(size_t)m * n + a] && masks[(size_t)m * n + b])) continue;
    int* par = parent + (size_t)m * n;
    const int ra = find_root(par, a), rb = find_root(par, b);
    if (ra == rb) continue;
    const int hi = ra > rb ? ra : rb, lo = ra > rb ? rb : ra;
    atomicMin(par + hi, lo);
    any = true;
  }
  if (any) *changed = 1;
}

__global__ void cc_jump_kernel(int* parent, long long total, int n) {
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    int* par = parent + (x / n) * n;
    const int v = (int)(x % n);
    par[v] = find_root(par, v);
  }
}

// per mask: {components among the selected points, selected points, edges with both ends selected}
__global__ void cc_count_kernel(const int* __restrict__ parent, int n, long long E, const int32_t* __restrict__ ei,
                                const int32_t* __restrict__ ej, int n_masks, const unsigned char* __restrict__ masks,
                                unsigned long long* out) {
  const int m = blockIdx.y;
  unsigned long long comp = 0, pts = 0, edges = 0;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const bool sel = !masks || masks[(size_t)m * n + v];
    pts += sel ? 1 : 0;
    comp += (sel && parent[(size_t)m * n + v] == (int)v) ? 1 : 0;
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x)
    edges += (!masks || (masks[(size_t)m * n + ei[e]] && masks[(size_t)m * n + ej[e]])) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    comp += __shfl_down_sync(0xffffffffu, comp, o);
    pts += __shfl_down_sync(0xffffffffu, pts, o);
    edges += __shfl_down_sync(0xffffffffu, edges, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 3 * m, comp); atomicAdd(out + 3 * m + 1, pts); atomicAdd(out + 3 * m + 2, edges);   // integer sums: order-free
  }
}

}  // namespace
}  // namespace tl

extern "C" int topolow_components(int64_t n, int64_t n_edges, const int32_t* edge_i, const int32_t* edge_j, int32_t n_masks,
                                  const uint8_t* masks, int64_t* components_out, int64_t* points_out, int64_t* edges_out,
                                  int32_t device) {
  using namespace tl;
  if (n < 1 || n >= (1ll << 31) || n_edges < 0 || n_masks < 1 || !components_out || (n_edges > 0 && (!edge_i || !edge_j)))
    return TOPOLOW_ERR_BAD_ARG;
  for (int64_t e = 0; e < n_edges; ++e)
    if (edge_i[e] < 0 || edge_j[e] < 0 || edge_i[e] >= n || edge_j[e] >= n) return TOPOLOW_ERR_BAD_ARG;
  try {
    TL_CUDA(cudaSetDevice(device));
    cudaStream_t s = cudaStreamPerThread;
    const long long total = (long long)n_masks * n;
    AsyncBuf<int> d_parent((size_t)total, s), d_changed(1, s);
    AsyncBuf<int32_t> d_ei((size_t)n_edges, s), d_ej((size_t)n_edges, s);
    AsyncBuf<unsigned char> d_masks(masks ? (size_t)total : 1, s);
    AsyncBuf<unsigned long long> d_out((size_t)3 * n_masks, s);
    if (n_edges > 0) {
      TL_CUDA(cudaMemcpyAsync(d_ei, edge_i, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      TL_CUDA(cudaMemcpyAsync(d_ej, edge_j, n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    if (masks) TL_CUDA(cudaMemcpyAsync(d_masks, masks, (size_t)total, cudaMemcpyHostToDevice, s));
    const unsigned char* mk = masks ? (const unsigned char*)d_masks : nullptr;
    const int blocks = 148 * 8, threads = 256;
    cc_init_kernel<<<blocks, threads, 0, s>>>(d_parent, total, (int)n);
    for (int pass = 0; n_edges > 0 && pass < 64; ++pass) {   // 64 passes: far beyond the O(log n) a graph needs
      int changed = 0;
      TL_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), s));
      cc_hook_kernel<<<blocks, threads, 0, s>>>(d_parent, (int)n, n_edges, d_ei, d_ej, n_masks, mk, d_changed);
      cc_jump_kernel<<<blocks, threads, 0, s>>>(d_parent, total, (int)n);
      TL_CUDA(cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, s));
      TL_CUDA(cudaStreamSynchronize(s));
      if (!changed) break;
    }
    TL_CUDA(cudaMemsetAsync(d_out, 0, (size_t)3 * n_masks * sizeof(unsigned long long), s));
    cc_count_kernel<<<dim3(64, (unsigned)n_masks), threads, 0, s>>>(d_parent, (int)n, n_edges, d_ei, d_ej, n_masks, mk, d_out);
    TL_CUDA(cudaGetLastError());
    std::vector<unsigned long long> out((size_t)3 * n_masks);
    TL_CUDA(cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    TL_CUDA(cudaStreamSynchronize(s));
    for (int m = 0; m < n_masks; ++m) {
      components_out[m] = (int64_t)out[3 * m];
      if (points_out) points_out[m] = (int64_t)out[3 * m + 1];
      if (edges_out) edges_out[m] = (int64_t)out[3 * m + 2];
    }
    return TOPOLOW_OK;
  } catch (const CudaError&) {
    cudaGetLastError();
    return TOPOLOW_ERR_CUDA;
  }
}
