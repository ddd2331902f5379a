// topolow_b200/csrc/tilepass_launch.h - host-visible launchers of the production kernel.
#pragma once
#include "tilepass.cuh"

namespace tl {

constexpr int kMaxDim = 16;

// Dynamic shared memory of one CTA: 2W tiles + W (target table + masks) + 2W tile ids + W flags.
inline size_t tile_smem_bytes(int D, int W, size_t real_size) {
  const size_t tile = (size_t)(D + 1) * kRow * real_size;
  return 2 * W * tile + (size_t)W * (kTableReals * real_size + kTableMasks * 4) + (size_t)3 * W * 4;
}

// Launch `n_iters` iterations (cooperative when geo.G > 1).  Throws CudaError.
void launch_tile_f32(const TileDev<float>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream);
void launch_tile_f64(const TileDev<double>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream);
// Largest CTA count of that instantiation that can be co-resident on the current device.
int max_coresident_f32(int D, int W);
int max_coresident_f64(int D, int W);

}  // namespace tl
