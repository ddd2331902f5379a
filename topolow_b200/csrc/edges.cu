// topolow_b200/csrc/edges.cu
//
// Device-side construction of the bucketed edge records of one fit: the COO edge list the R
// caller builds (R/core.R:383-402: edge_i, edge_j, edge_dist, edge_thresh) becomes 16-byte
// records sorted by (tile of the lower slot, tile of the higher slot), plus the bucket offsets.
// Counting sort with atomics, followed by a rank sort inside every bucket so that the result -
// and with it the summation order of the edge MAE - is the same on every run.
#include "edges.h"

#include <vector>

namespace tl {
namespace {

__global__ void key_kernel(const int32_t* __restrict__ ei, const int32_t* __restrict__ ej, long long E, long long n,
                           const int32_t* __restrict__ slot_of_point, int T, uint32_t kTile, uint32_t* __restrict__ keys,
                           uint32_t* __restrict__ counts, int* __restrict__ bad) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long a = ei[e], b = ej[e];
    if (a < 0 || b < 0 || a >= n || b >= n || a == b) { *bad = 1; keys[e] = 0; continue; }
    uint32_t sa = (uint32_t)slot_of_point[a], sb = (uint32_t)slot_of_point[b];
    if (sa > sb) { const uint32_t t = sa; sa = sb; sb = t; }
    const uint32_t key = (sa / kTile) * (uint32_t)T + (sb / kTile);
    keys[e] = key;
    atomicAdd(&counts[key + 1], 1u);
  }
}

// Inclusive scan of `data` in place, 1024 elements per block; block totals to `sums`.
__global__ void scan_block_kernel(uint32_t* __restrict__ data, size_t count, uint32_t* __restrict__ sums) {
  __shared__ uint32_t sh[1024];
  const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  sh[threadIdx.x] = i < count ? data[i] : 0u;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  if (i < count) data[i] = sh[threadIdx.x];
  if (threadIdx.x == 1023) sums[blockIdx.x] = sh[1023];
}
__global__ void scan_add_kernel(uint32_t* __restrict__ data, size_t count, const uint32_t* __restrict__ block_prefix) {
  const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < count && blockIdx.x > 0) data[i] += block_prefix[blockIdx.x - 1];
}

__global__ void scatter_kernel(const int32_t* __restrict__ ei, const int32_t* __restrict__ ej,
                               const double* __restrict__ dist, const int32_t* __restrict__ thr, long long E,
                               const int32_t* __restrict__ slot_of_point, const uint32_t* __restrict__ keys,
                               uint32_t* __restrict__ cursor, EdgeRec* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    uint32_t sa = (uint32_t)slot_of_point[ei[e]], sb = (uint32_t)slot_of_point[ej[e]];
    if (sa > sb) { const uint32_t t = sa; sa = sb; sb = t; }
    const int t = thr[e];
    EdgeRec r;
    r.target = dist[e];
    r.slot_lo = sa;
    r.slot_hi_type = sb | ((uint32_t)(t == 0 ? 0 : (t == 1 ? 1 : 2))  /* src/optimization.cpp:237-243 */ << 30);
    out[atomicAdd(&cursor[keys[e]], 1u)] = r;
  }
}

// One warp per bucket: out[rank] = in[e], rank = number of records of the bucket that precede e in
// (slot_lo, slot_hi, target bits, type) order.  A pair listed twice (also as (j, i)) is legal input: records
// with equal keys are ordered by their remaining bits, fully equal records by their position, so the ranks
// are a permutation and the result does not depend on the order the scatter left.
__global__ void bucket_sort_kernel(const EdgeRec* __restrict__ in, EdgeRec* __restrict__ out,
                                   const uint32_t* __restrict__ off, size_t nkeys) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t key = warp; key < nkeys; key += nwarps) {
    const uint32_t b = off[key], e = off[key + 1];
    for (uint32_t x = b + lane; x < e; x += 32) {
      const EdgeRec r = in[x];
      const unsigned long long mine = ((unsigned long long)r.slot_lo << 32) | (r.slot_hi_type & 0x3fffffffu);
      const unsigned long long mine2 = ((unsigned long long)__double_as_longlong(r.target) << 2) ^ (r.slot_hi_type >> 30);
      uint32_t rank = 0;
      for (uint32_t y = b; y < e; ++y) {
        const EdgeRec o = in[y];
        const unsigned long long other = ((unsigned long long)o.slot_lo << 32) | (o.slot_hi_type & 0x3fffffffu);
        const unsigned long long other2 = ((unsigned long long)__double_as_longlong(o.target) << 2) ^ (o.slot_hi_type >> 30);
        const bool before = other < mine || (other == mine && (other2 < mine2 || (other2 == mine2 && y < x)));
        rank += before ? 1u : 0u;
      }
      out[b + rank] = r;
    }
  }
}

}  // namespace

void build_buckets(const topolow_problem& pb, const std::vector<int32_t>& slot_of_point, int T, int tile_points,
                   cudaStream_t stream,
                   EdgeRec** edges_out, uint32_t** bucket_off_out) {
  const long long E = pb.n_edges;
  const size_t nkeys = (size_t)T * T;
  // the two results live as long as the plan; everything else is stream-ordered scratch
  AsyncBuf<uint32_t> d_off(nkeys + 1, stream);   // released to the caller: free with pool_free()
  AsyncBuf<EdgeRec> d_out(E, stream);
  AsyncBuf<int32_t> d_ei(E, stream), d_ej(E, stream), d_thr(E, stream), d_slot(slot_of_point.size(), stream);
  AsyncBuf<double> d_dist(E, stream);
  AsyncBuf<uint32_t> d_keys(E, stream), d_cur(nkeys + 1, stream);
  const size_t nblk = (nkeys + 1 + 1023) / 1024;
  AsyncBuf<uint32_t> d_sums(nblk, stream);
  AsyncBuf<int> d_bad(1, stream);
  AsyncBuf<EdgeRec> d_tmp(E, stream);
  PhaseTimer pt(stream);
  pt.mark("edges: allocations");
  {
    TL_CUDA(cudaMemcpyAsync(d_slot, slot_of_point.data(), slot_of_point.size() * 4, cudaMemcpyHostToDevice, stream));
    if (E > 0) {
      TL_CUDA(cudaMemcpyAsync(d_ei, pb.edge_i, E * 4, cudaMemcpyHostToDevice, stream));
      TL_CUDA(cudaMemcpyAsync(d_ej, pb.edge_j, E * 4, cudaMemcpyHostToDevice, stream));
      TL_CUDA(cudaMemcpyAsync(d_dist, pb.edge_dist, E * 8, cudaMemcpyHostToDevice, stream));
      TL_CUDA(cudaMemcpyAsync(d_thr, pb.edge_thresh, E * 4, cudaMemcpyHostToDevice, stream));
    }
    pt.mark("edges: host to device");
    TL_CUDA(cudaMemsetAsync(d_off, 0, (nkeys + 1) * 4, stream));
    TL_CUDA(cudaMemsetAsync(d_bad, 0, 4, stream));
    const int blocks = 148 * 8, threads = 256;
    if (E > 0) {
      key_kernel<<<blocks, threads, 0, stream>>>(d_ei, d_ej, E, pb.n, d_slot, T, (uint32_t)tile_points, d_keys, d_off, d_bad);
      TL_CUDA(cudaGetLastError());
    }
    // exclusive offsets: counts were written at key + 1, so an inclusive scan yields them
    scan_block_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(d_off, nkeys + 1, d_sums);
    TL_CUDA(cudaGetLastError());
    std::vector<uint32_t> sums(nblk);
    int bad = 0;
    TL_CUDA(cudaMemcpyAsync(sums.data(), d_sums, nblk * 4, cudaMemcpyDeviceToHost, stream));
    TL_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, stream));
    TL_CUDA(cudaStreamSynchronize(stream));
    pt.mark("edges: keys + scan");
    if (bad) throw std::invalid_argument("edge index out of range");
    for (size_t i = 1; i < nblk; ++i) sums[i] += sums[i - 1];
    TL_CUDA(cudaMemcpyAsync(d_sums, sums.data(), nblk * 4, cudaMemcpyHostToDevice, stream));
    scan_add_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(d_off, nkeys + 1, d_sums);
    TL_CUDA(cudaGetLastError());
    if (E > 0) {
      TL_CUDA(cudaMemcpyAsync(d_cur, d_off, (nkeys + 1) * 4, cudaMemcpyDeviceToDevice, stream));
      scatter_kernel<<<blocks, threads, 0, stream>>>(d_ei, d_ej, d_dist, d_thr, E, d_slot, d_keys, d_cur, d_tmp);
      TL_CUDA(cudaGetLastError());
      pt.mark("edges: scatter");
      bucket_sort_kernel<<<blocks, threads, 0, stream>>>(d_tmp, d_out, d_off, nkeys);
      TL_CUDA(cudaGetLastError());
    }
    TL_CUDA(cudaStreamSynchronize(stream));
    pt.mark("edges: sort inside buckets");
  }
  *edges_out = d_out.release();
  *bucket_off_out = d_off.release();
}

}  // namespace tl
