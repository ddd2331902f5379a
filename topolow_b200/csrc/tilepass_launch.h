// topolow_b200/csrc/tilepass_launch.h - host-visible launchers of the production kernel.
#pragma once
#include "tiledev.h"

namespace tl {

constexpr int kMaxDim = 16;
constexpr int kMaxP = 3;

// Warps per CTA the kernel is built for, by precision (0 = FP32, 1 = exact FP64) and points per lane.
inline int tile_max_warps(int precision, int P) {
  const int f32[4] = {0, 16, 8, 4}, f64[4] = {0, 8, 4, 2};
  return precision ? f64[P] : f32[P];
}
// Dynamic shared memory of one CTA: 2W tiles + W (target tables + masks) + 2W tile ids + W flags +
// W staging areas of 64 P edge records (16-byte aligned).
inline size_t tile_smem_bytes(int D, int W, size_t real_size, int P) {
  const size_t tile = (size_t)(D + 1) * (32 * P + 1) * real_size;
  return 2 * W * tile + (size_t)W * ((size_t)P * P * 1024 * real_size + (size_t)P * P * 64 * 4) + (size_t)3 * W * 4 +
         16 + (size_t)W * 64 * P * 16;
}

// Shared memory a launch may use besides the kernel's static ~1.5 KB, and the lookup table of a geometry.
constexpr size_t kSmemLaunchLimit = 227 * 1024 - 2048;
inline size_t perm_table_bytes(const Geometry& g) { return 4 * (size_t)perm_table_entries(g); }
// Decide whether the geometry's lookup table fits next to `base` bytes; sets g.table, returns the total.
inline size_t with_perm_table(Geometry& g, size_t base) {
  g.table = base + perm_table_bytes(g) <= kSmemLaunchLimit ? 1 : 0;
  return g.table ? base + perm_table_bytes(g) : base;
}

// Launch `n_iters` iterations (cooperative when geo.G > 1); geo.P selects the tile size.  Throws CudaError.
void launch_tile_f32(const TileDev<float>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream);
void launch_tile_f64(const TileDev<double>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream);
// Many fits in one launch, one CTA each, 32- or 64-point tiles (P = 1 or 2): d_jobs is a device array of
// n_jobs entries with equal D, W and P.  n_jobs == 0 only loads the kernel.
// smem = dynamic shared memory of every CTA (tile_smem_bytes + the largest lookup table of the jobs).
void launch_tile_batch_f32(int D, int P, const BatchJob<float>* d_jobs, int n_jobs, int W, size_t smem, cudaStream_t stream);
void launch_tile_batch_f64(int D, int P, const BatchJob<double>* d_jobs, int n_jobs, int W, size_t smem, cudaStream_t stream);
// Largest CTA count of that instantiation that can be co-resident on the current device.
int max_coresident_f32(int D, int W, int P);
int max_coresident_f64(int D, int W, int P);

}  // namespace tl
