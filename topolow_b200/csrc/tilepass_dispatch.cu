// topolow_b200/csrc/tilepass_dispatch.cu - picks the tile-size instantiation named by geo.P.
#include "tilepass_launch.h"

namespace tl {
#define TL_DECL(SUF, REAL, P)                                                                              \
  void launch_tile_##SUF##_p##P(const TileDev<REAL>&, const Geometry&, const FitParams&, int, volatile int*, \
                                cudaStream_t);                                                             \
  int max_coresident_##SUF##_p##P(int, int);
TL_DECL(f32, float, 1) TL_DECL(f32, float, 2) TL_DECL(f32, float, 3)
TL_DECL(f64, double, 1) TL_DECL(f64, double, 2) TL_DECL(f64, double, 3)
#undef TL_DECL

void launch_tile_f32(const TileDev<float>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream) {
  if (geo.P == 1) launch_tile_f32_p1(dv, geo, prm, n_iters, host_flag, stream);
  else if (geo.P == 2) launch_tile_f32_p2(dv, geo, prm, n_iters, host_flag, stream);
  else launch_tile_f32_p3(dv, geo, prm, n_iters, host_flag, stream);
}
void launch_tile_f64(const TileDev<double>& dv, const Geometry& geo, const FitParams& prm, int n_iters,
                     volatile int* host_flag, cudaStream_t stream) {
  if (geo.P == 1) launch_tile_f64_p1(dv, geo, prm, n_iters, host_flag, stream);
  else if (geo.P == 2) launch_tile_f64_p2(dv, geo, prm, n_iters, host_flag, stream);
  else launch_tile_f64_p3(dv, geo, prm, n_iters, host_flag, stream);
}
#define TL_DECLB(SUF, REAL, P) void launch_tile_batch_##SUF##_p##P(int, const BatchJob<REAL>*, int, int, size_t, cudaStream_t);
TL_DECLB(f32, float, 1) TL_DECLB(f32, float, 2) TL_DECLB(f64, double, 1) TL_DECLB(f64, double, 2)
#undef TL_DECLB
void launch_tile_batch_f32(int D, int P, const BatchJob<float>* d_jobs, int n_jobs, int W, size_t smem, cudaStream_t stream) {
  if (P == 1) launch_tile_batch_f32_p1(D, d_jobs, n_jobs, W, smem, stream);
  else launch_tile_batch_f32_p2(D, d_jobs, n_jobs, W, smem, stream);
}
void launch_tile_batch_f64(int D, int P, const BatchJob<double>* d_jobs, int n_jobs, int W, size_t smem, cudaStream_t stream) {
  if (P == 1) launch_tile_batch_f64_p1(D, d_jobs, n_jobs, W, smem, stream);
  else launch_tile_batch_f64_p2(D, d_jobs, n_jobs, W, smem, stream);
}
int max_coresident_f32(int D, int W, int P) {
  return P == 1 ? max_coresident_f32_p1(D, W) : (P == 2 ? max_coresident_f32_p2(D, W) : max_coresident_f32_p3(D, W));
}
int max_coresident_f64(int D, int W, int P) {
  return P == 1 ? max_coresident_f64_p1(D, W) : (P == 2 ? max_coresident_f64_p2(D, W) : max_coresident_f64_p3(D, W));
}
}  // namespace tl
