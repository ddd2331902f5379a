// topolow_b200/csrc/rowblock.cu
//
// Row-block ("owner computes") mode of the embedding loop - the partitioning of one large map that
// SURVEY.md section 8e / BASELINE.json's north_star describe: the points are cut into G contiguous row
// blocks, GPU g owns the updates of its rows and computes them against a replica of ALL positions;
// within an iteration a point's own springs are applied one after another (Gauss-Seidel along the row),
// other points are read at their position of the iteration's start (Jacobi across points and across
// GPUs); the new rows are stored straight into every replica over NVLink (peer stores) and one
// flag exchange per iteration is the only synchronisation.  G = 1 is the same code on one GPU.
//
// One iteration on the replica P of all positions (reference: /root/reference/src/optimization.cpp):
//   repulse_kernel  R_i = -sum_{j != i} (P_j - P_i) * c / (2 (|P_j - P_i| + 0.01)^3) / (deg_i + 1)   (:269-281
//                   seen from i's side), a tiled one-sided N-body pass in packed FP32: a work item is
//                   256 own rows x one chunk of 2048 partners streamed through shared memory by cp.async.
//   spring_kernel   x = P_i, then for every measured pair (i, j) of row i in turn (:226-256, i's side):
//                   spring iff exact, or '>' and dist < target, or '<' and dist > target (:237-243);
//                   x -= (P_j - x) * 2k (target - dist) / (dist + 0.01) / (4 (deg_i + 1) + k)  (:246-253) and the
//                   repulsion R_i contains for this pair is taken back (a pair in spring state is not
//                   repelled); a satisfied threshold keeps its repulsion (:257-267).  One thread per row, records
//                   in degree-padded slices of 32 rows (SELL-32) so that a warp's record loads are one contiguous
//                   256-byte line.  Runs on a second stream NEXT TO the repulsion pass (both read only P).
//   combine_kernel  P'_i = x + R_i goes to every replica (peer stores), cooling (:289), one flag per peer.
//   mae_kernel      on check iterations: sum |target - dist| and count over the measured pairs that are exact
//                   or violated (:54-81) on P', FP64 sums, fixed summation tree; partial sums go to every replica.
//   ctl_kernel      cooling happened in the spring kernel (:289); three-way controller with best-state snapshot
//                   (:303-357, common.cuh::controller_check), finite check every 10 iterations (:359-361).
// Every unordered pair is visited once from each side per iteration and each side moves only its own
// endpoint, by what the reference's visit of the pair would move it.  Results do not depend on G (the
// summation trees are fixed), which is what the tests check; the scheme is compared with the reference's
// sequential loop statistically (tests/test_gpu_rowblock.py) and with its own CPU restatement
// (oracle/relaxed_oracle.cpp) numerically.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "plan.h"
#include "rowblock.h"

namespace tl {

std::vector<int32_t> random_permutation(int64_t n, uint64_t seed);   // plan.cu

namespace {

constexpr uint64_t kRowLayoutSeed = 0x726f77626c6f636bULL;
constexpr int kStageJ = 128;   // partners per shared-memory stage
constexpr int kStages = 3;
constexpr int kFlagStride = 32;   // 32-bit words between the flag words of two ranks (128 bytes)
constexpr unsigned long long kWaitNs = 20ull * 1000ull * 1000ull * 1000ull;

struct RowDev {
  int n, slots, D, Dp, G, rank, row0, rows, chunks;
  int rparts;                        // partial-sum entries per row in rpart: chunks (2 x chunks for the tensor-core pass)
  int adaptive;                      // 1: the repulsion form is chosen per iteration on the device (counters[6], set by image_tc_kernel)
  int series;                        // tensor form: weights by the one-MUFU series where the probe allows it (adaptive) / always (not adaptive)
  int probe_k;                       // sampled partners per row of the near-pair probe
  unsigned probe_limit;              // tensor form while the probe finds at most this many near pairs
  int no_wait;                       // measurement only (TOPOLOW_IGNORE_PEERS): one rank of a sharded map timed without its peers
  unsigned long long cap_rows;       // rows of one position buffer (slots rounded up to a chunk)
  float* pos[kMaxShards];            // replica q: [2][cap_rows][Dp]
  double* red[kMaxShards];           // replica q: [slots / 128][4]  {sum |err|, count, non-finite, -}
  unsigned* flags[kMaxShards];       // replica q: [kMaxShards][kFlagStride], word r = epoch reached by rank r
  float* best;                       // [cap_rows][Dp]
  const float* dp1;                  // [slots] degree + 1 (0 = padding row)
  float* rpart;                      // [chunks][rows][Dp] repulsion sums per partner chunk
  float* xs;                         // [rows][Dp] positions of the own rows after the spring walk (before the repulsion is added)
  float* img;                        // [cap_rows][Img<H>::kStride] partner image of the iteration: -2 (P_j - centre) | |P_j - centre|^2 / 2 twice
  float* hmax;                       // [2] max over all points of |P_j - centre|^2 / 2, per position buffer
  float centre[16];                  // centroid of the initial positions (fixed for the fit, the same on every rank)
  const uint2* recs;                 // spring records of the own rows, SELL-32: {partner | type << 30, target}
  const unsigned long long* soff;    // [rows / 32] first record of a slice
  const int* swidth;                 // [rows / 32] records per row of a slice
  const uint2* mrecs;                // the records this rank counts in the MAE (every pair once over all ranks)
  const unsigned long long* moff;
  const int* mwidth;
  FitState* state;
  double* trace;
  unsigned* counters;                // [0..2] tickets of the "last CTA" of a launch / work items; [3], [4] near pairs the probe found (series
                                     // form / plain), [5] its ticket, [6] form of this iteration (0 FP32, 1 tensor, 2 tensor with the series
                                     // weights), [7] iterations run in a tensor form
  volatile int* host_flag;           // mapped: [0] stop, [1] iterations done
  unsigned long long pairs_per_iter;
  unsigned long long seed;
};

typedef void (*RepulseFn)(RowDev, int, unsigned);

// Adaptive runs launch both repulsion kernels every iteration; the one whose form was not chosen returns at once.
TL_D bool form_is_tensor(const RowDev& dv) { return __ldcg(&dv.counters[6]) != 0u; }

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
TL_D float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
TL_D float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
TL_D unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
TL_D void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TL_D unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
TL_D void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
TL_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> TL_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Rows are stored with a stride of kStride<H> floats (a multiple of 4: every row starts on a 16-byte boundary,
// odd H leaves two pad floats that are never computed on).
template <int H> struct Row { static constexpr int kStride = (2 * H + 3) / 4 * 4; };

// H packed pairs of one point from (shared or global) memory: 16-byte loads, one 8-byte tail when H is odd.
template <int H>
TL_D void ld_point(const float* __restrict__ p, float2 (&v)[H]) {
#pragma unroll
  for (int k = 0; k < H / 2; ++k) {
    const float4 t = reinterpret_cast<const float4*>(p)[k];
    v[2 * k] = make_float2(t.x, t.y); v[2 * k + 1] = make_float2(t.z, t.w);
  }
  if constexpr (H % 2 == 1) v[H - 1] = reinterpret_cast<const float2*>(p)[H - 1];
}
template <int H>
TL_D void st_point(float* __restrict__ p, const float2 (&v)[H]) {
#pragma unroll
  for (int k = 0; k < H / 2; ++k)
    reinterpret_cast<float4*>(p)[k] = make_float4(v[2 * k].x, v[2 * k].y, v[2 * k + 1].x, v[2 * k + 1].y);
  if constexpr (H % 2 == 1) reinterpret_cast<float2*>(p)[H - 1] = v[H - 1];
}

// |q + np|^2 (np = -p); delta is kept for the caller.  kChains = 2 halves the dependent FMA chain (the
// spring walk is one long dependency chain), 1 saves the packed add that joins the chains (the repulsion
// pass has enough independent interactions in flight and is bound by the FMA pipe).
template <int H, int kChains = 2>
TL_D float dist2(const float2 (&np)[H], const float2 (&q)[H], float2 (&dl)[H]) {
#pragma unroll
  for (int k = 0; k < H; ++k) dl[k] = __fadd2_rn(q[k], np[k]);
  float2 s0 = __fmul2_rn(dl[0], dl[0]);
  if constexpr (kChains == 1 || H == 1) {
#pragma unroll
    for (int k = 1; k < H; ++k) s0 = __ffma2_rn(dl[k], dl[k], s0);
  } else {
    float2 s1 = __fmul2_rn(dl[1], dl[1]);
#pragma unroll
    for (int k = 2; k < H; ++k) {
      if (k & 1) s1 = __ffma2_rn(dl[k], dl[k], s1);
      else s0 = __ffma2_rn(dl[k], dl[k], s0);
    }
    s0 = __fadd2_rn(s0, s1);
  }
  return s0.x + s0.y;
}

TL_D float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
TL_D float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// One-sided repulsion of partner q on the point -np: acc += (q - p) / (|q - p| + 0.01)^3.
// kPow: the cube and its reciprocal as 2^(-3 log2(ds)) on the special-function unit (two MUFU + one multiply)
// instead of two multiplies + MUFU.RCP: one FMA-pipe slot less in a loop that is bound by that pipe.
template <int H, bool kPow = false>
TL_D void repel(const float2 (&np)[H], const float2 (&q)[H], float2 (&acc)[H]) {
  float2 dl[H];
  const float d2 = dist2<H, 1>(np, q, dl);
  const float ds = sqrt_approx(d2) + 0.01f;
  const float w = kPow ? ex2_approx(-3.0f * lg2_approx(ds)) : rcp_approx(ds * ds * ds);
  const float2 ww = make_float2(w, w);
#pragma unroll
  for (int k = 0; k < H; ++k) acc[k] = __ffma2_rn(dl[k], ww, acc[k]);
}

// Every thread of the CTA calls this; returns false when the peers did not arrive in time (the fit is
// then stopped with an error instead of hanging the device).
TL_D bool wait_epoch(const RowDev& dv, unsigned epoch) {
  if (dv.G <= 1 || epoch == 0 || dv.no_wait) return true;
  __shared__ int ok_s;
  if (threadIdx.x < 32) {
    const unsigned* f = dv.flags[dv.rank] + (size_t)(threadIdx.x < (unsigned)dv.G ? threadIdx.x : dv.rank) * kFlagStride;
    const unsigned long long t0 = global_ns();
    bool ok = true;
    for (;;) {
      const unsigned v = ld_acquire_sys(f);
      if (__all_sync(0xffffffffu, v >= epoch)) break;
      if (global_ns() - t0 > kWaitNs) { ok = false; break; }
      __nanosleep(200);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) ok_s = ok ? 1 : 0;
  }
  __syncthreads();
  const bool ok = ok_s != 0;
  __syncthreads();
  return ok;
}
TL_D void peer_timeout(const RowDev& dv) {   // one thread
  dv.state->status = 3; dv.state->stop = 1;
  if (dv.host_flag) { __threadfence_system(); dv.host_flag[0] = 1; }
}
// One thread, after a __threadfence_system() that covers the data: tell every replica that this rank reached `epoch`.
TL_D void signal_epoch(const RowDev& dv, unsigned epoch) {
  for (int q = 0; q < dv.G; ++q) st_release_sys(dv.flags[q] + (size_t)dv.rank * kFlagStride, epoch);
}
// All threads call; true in every thread of the CTA that finished last.
TL_D bool last_cta(unsigned* counter) {
  __shared__ int last_s;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    last_s = (t == gridDim.x - 1) ? 1 : 0;
    if (last_s) { *counter = 0u; __threadfence_system(); }
  }
  __syncthreads();
  return last_s != 0;
}

// ---------------------------------------------------------------------------------------------------
// repulsion: one work item = 256 own rows (R per thread) x one chunk of <= 2048 partners
// ---------------------------------------------------------------------------------------------------
// Items are handed out by an atomic counter (reset by the spring kernel's last CTA): the partial sums of an
// item do not depend on which CTA computed it, so the result is the same whatever the residency pattern.
template <int H, int R, int U, int MAXR, bool kPow>
__global__ void __maxnreg__(MAXR) repulse_kernel(RowDev dv, int cur, unsigned epoch) {
  extern __shared__ float4 smem4[];
  __shared__ long long item_s;
  float* sm = reinterpret_cast<float*>(smem4);
  constexpr int T = kRowTile / R;
  constexpr int Dp = Row<H>::kStride;
  constexpr int kStageFloats = kStageJ * Dp;
  if (__ldcg(&dv.state->stop)) return;
  if (dv.adaptive && form_is_tensor(dv)) return;
  if (!wait_epoch(dv, epoch)) { if (blockIdx.x == 0 && threadIdx.x == 0) peer_timeout(dv); return; }
  const int tid = threadIdx.x;
  const float* __restrict__ P = dv.pos[dv.rank] + (size_t)cur * dv.cap_rows * Dp;
  const int tiles = dv.rows / kRowTile;
  const long long items = (long long)tiles * dv.chunks;
  for (;;) {
    if (tid == 0) item_s = (long long)atomicAdd(&dv.counters[2], 1u);
    __syncthreads();
    const long long item = item_s;
    if (item >= items) break;
    const int c = (int)(item / tiles), tile = (int)(item % tiles);
    const int lrow0 = tile * kRowTile;
    float2 np[R][H], acc[R][H];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 p[H];
      ld_point<H>(P + (size_t)(dv.row0 + lrow0 + tid + r * T) * Dp, p);
#pragma unroll
      for (int k = 0; k < H; ++k) { np[r][k] = make_float2(-p[k].x, -p[k].y); acc[r][k] = make_float2(0.f, 0.f); }
    }
    const int j0 = c * kChunk;
    const int cnt = min(kChunk, dv.n - j0);
    const int nst = (cnt + kStageJ - 1) / kStageJ;
    const float* __restrict__ src = P + (size_t)j0 * Dp;
    auto prefetch = [&](int s) {
      if (s < nst) {
        const float4* g = reinterpret_cast<const float4*>(src + (size_t)s * kStageFloats);
        float4* d = reinterpret_cast<float4*>(sm + (size_t)(s % kStages) * kStageFloats);
        for (int x = tid; x < kStageFloats / 4; x += T) cp_async16(d + x, g + x);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) prefetch(s);
    for (int s = 0; s < nst; ++s) {
      cp_async_wait<kStages - 2>();
      __syncthreads();
      prefetch(s + kStages - 1);
      const float* __restrict__ q_s = sm + (size_t)(s % kStages) * kStageFloats;
      const int m = min(kStageJ, cnt - s * kStageJ);
#pragma unroll U
      for (int j = 0; j < m; ++j) {
        float2 q[H];
        ld_point<H>(q_s + j * Dp, q);
#pragma unroll
        for (int r = 0; r < R; ++r) repel<H, kPow>(np[r], q, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      st_point<H>(dv.rpart + ((size_t)c * dv.rows + lrow0 + tid + r * T) * Dp, acc[r]);
    __syncthreads();   // the stages are refilled by the next item's prologue; item_s is rewritten
  }
}

// ---------------------------------------------------------------------------------------------------
// repulsion, inner-product form (ndim >= 5): |q - p|^2 = |p|^2 + |q|^2 - 2 p.q
// ---------------------------------------------------------------------------------------------------
// The difference form above spends 3 packed FMA-pipe instructions per coordinate pair and interaction
// (q - p, its square, the accumulation of w (q - p)); the pass is bound by that pipe (ncu: 82 % busy).  Written
// with inner products an interaction needs 2: p.q~ (q~ = -2 (q - centre), its half norm rides in the accumulator's
// start value) and A += w q~, plus W += w; the sum over the partners comes out as -A / 2 - p W at the end of the
// item.  image_kernel prepares q~ and the half norms once per iteration (6 MB read, 8 MB written).
// Cancellation: the rounding error of the inner-product distance is about (H + 1) 2^-24 (|p|^2 + |q|^2), harmless for
// a pair that is far apart and fatal for one that is not.  Pairs with d^2 < thr = (H + 1) 2^-13 (|p|^2 / 2 + max_j |q_j|^2 / 2)
// ("near": relative error of d^2 could exceed 1e-3; a handful per point in 5 dimensions, none in 16, always the
// point itself) get weight 0 in the main loop, raise a flag, and the stage is walked again for them in the
// difference form into a thread-private accumulator in shared memory - the same arithmetic as repulse_kernel.
template <int H> struct Img { static constexpr int kStride = (2 * H + 2 + 3) / 4 * 4; };

template <int NF4>
TL_D float f4_elem(const float4 (&t)[NF4], int e) {
  const float4& x = t[e >> 2];
  switch (e & 3) { case 0: return x.x; case 1: return x.y; case 2: return x.z; default: return x.w; }
}
// image row: H packed pairs of q~ and the half norm (in both lanes of hh)
template <int H>
TL_D void ld_image(const float* __restrict__ p, float2 (&v)[H], float2& hh) {
  constexpr int NF4 = Img<H>::kStride / 4;
  float4 t[NF4];
#pragma unroll
  for (int k = 0; k < NF4; ++k) t[k] = reinterpret_cast<const float4*>(p)[k];
#pragma unroll
  for (int k = 0; k < H; ++k) v[k] = make_float2(f4_elem<NF4>(t, 2 * k), f4_elem<NF4>(t, 2 * k + 1));
  hh = make_float2(f4_elem<NF4>(t, 2 * H), f4_elem<NF4>(t, 2 * H + 1));
}
// d^2 by inner product and the near test.  Explicit roundings: the main loop and the second walk must agree bit for bit.
template <int H>
TL_D bool near_test(const float2 (&p)[H], float sp, float thr, const float2 (&q)[H], float2 hh, float& d2) {
  float2 s = __ffma2_rn(p[0], q[0], hh);
#pragma unroll
  for (int k = 1; k < H; ++k) s = __ffma2_rn(p[k], q[k], s);
  d2 = __fadd_rn(__fadd_rn(s.x, s.y), sp);
  return d2 < thr;
}
// kSeries: (dist + 0.01)^-3 = (r (1 - x + x^2 - x^3))^3 (1 + O(x^4)) with r = 1 / dist (one MUFU.RSQ) and x = 0.01 r -
// seven FMA-pipe instructions instead of three MUFU (sqrt, lg2, ex2), each of which holds the quarter's
// special-function unit for 8 cycles.  Only pairs that pass the near test get here: the threshold is never below
// kSeriesFloor = 0.1 (dist > 0.316, x < 0.0316, x^4 < 1e-6), the others take the exact formula in the second walk.
constexpr float kSeriesFloor = 0.1f;
TL_D float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int H, bool kSeries>
TL_D bool repel_dot(const float2 (&p)[H], float sp, float thr, const float2 (&q)[H], float2 hh, float2 (&acc)[H], float& W) {
  float d2;
  const bool nr = near_test<H>(p, sp, thr, q, hh, d2);
  float w;
  if constexpr (kSeries) {
    const float r = rsqrt_approx(d2);
    const float a = fmaf(-0.01f, r, 1.0f);
    const float b = (r * r) * 1e-4f;
    const float u = r * fmaf(b, a, a);
    w = u * u * u;
  } else {
    const float ds = sqrt_approx(d2) + 0.01f;
    w = ex2_approx(-3.0f * lg2_approx(ds));
  }
  w = nr ? 0.f : w;                      // also swallows the NaN of a slightly negative d2
  const float2 ww = make_float2(w, w);
#pragma unroll
  for (int k = 0; k < H; ++k) acc[k] = __ffma2_rn(q[k], ww, acc[k]);
  W += w;
  return nr;
}

template <int H>
__global__ void __launch_bounds__(kBlockRows) image_kernel(RowDev dv, int cur, unsigned epoch) {
  constexpr int Dp = Row<H>::kStride, Dq = Img<H>::kStride;
  if (__ldcg(&dv.state->stop)) return;
  if (!wait_epoch(dv, epoch)) { if (blockIdx.x == 0 && threadIdx.x == 0) peer_timeout(dv); return; }
  const int row = blockIdx.x * kBlockRows + threadIdx.x;
  if (row == 0) dv.hmax[cur ^ 1] = 0.f;          // the other buffer's maximum: its readers (the previous iteration) are done
  const float* __restrict__ P = dv.pos[dv.rank] + (size_t)cur * dv.cap_rows * Dp;
  float2 p[H];
  ld_point<H>(P + (size_t)row * Dp, p);
  float4 o[Dq / 4];
  float* of = reinterpret_cast<float*>(o);
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < H; ++k) {
    const float2 dl = __fadd2_rn(p[k], make_float2(-dv.centre[2 * k], -dv.centre[2 * k + 1]));
    s = __ffma2_rn(dl, dl, s);
    of[2 * k] = -2.0f * dl.x; of[2 * k + 1] = -2.0f * dl.y;
  }
  float h = 0.5f * __fadd_rn(s.x, s.y);
  if (row >= dv.n) {                              // padding rows: never partners; keep them finite
    h = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * H; ++k) of[k] = 0.f;
  }
#pragma unroll
  for (int k = 2 * H; k < Dq; ++k) of[k] = h;
  float4* out = reinterpret_cast<float4*>(dv.img + (size_t)row * Dq);
#pragma unroll
  for (int k = 0; k < Dq / 4; ++k) out[k] = o[k];
  float m = h;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(dv.hmax + cur), __float_as_int(m));   // m >= 0: integer order = float order
}

template <int H, int R, int U, int MAXR, int SJ = kStageJ, bool kSeries = false>
__global__ void __maxnreg__(MAXR) repulse_dot_kernel(RowDev dv, int cur, unsigned epoch) {
  extern __shared__ float4 smem4[];
  __shared__ long long item_s;
  float* sm = reinterpret_cast<float*>(smem4);
  constexpr int T = kRowTile / R;
  constexpr int Dp = Row<H>::kStride, Dq = Img<H>::kStride;
  constexpr int kStageFloats = SJ * Dq;
  float2* fix_s = reinterpret_cast<float2*>(sm + kStages * kStageFloats);   // [R * H][T], written by the second walk only
  if (__ldcg(&dv.state->stop)) return;
  (void)epoch;                                     // image_kernel of this iteration waited for the peers
  const int tid = threadIdx.x;
  const float* __restrict__ I = dv.img;
  const float hmax = __ldcg(dv.hmax + cur);
  constexpr float kTau = (float)(H + 1) / 8192.0f;
  const int tiles = dv.rows / kRowTile;
  const long long items = (long long)tiles * dv.chunks;
  for (;;) {
    if (tid == 0) item_s = (long long)atomicAdd(&dv.counters[2], 1u);
    __syncthreads();
    const long long item = item_s;
    if (item >= items) break;
    const int c = (int)(item / tiles), tile = (int)(item % tiles);
    const int lrow0 = tile * kRowTile;
    float2 p[R][H], acc[R][H];
    float sp[R], thr[R], W[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 qt[H], hh;
      ld_image<H>(I + (size_t)(dv.row0 + lrow0 + tid + r * T) * Dq, qt, hh);
#pragma unroll
      for (int k = 0; k < H; ++k) { p[r][k] = make_float2(-0.5f * qt[k].x, -0.5f * qt[k].y); acc[r][k] = make_float2(0.f, 0.f); }
      sp[r] = 2.0f * hh.x; thr[r] = kTau * (hh.x + hmax); W[r] = 0.f;
      if constexpr (kSeries) thr[r] = fmaxf(thr[r], kSeriesFloor);
    }
    bool fixed = false;
    const int j0 = c * kChunk;
    const int cnt = min(kChunk, dv.n - j0);
    const int nst = (cnt + SJ - 1) / SJ;
    const float* __restrict__ src = I + (size_t)j0 * Dq;
    auto prefetch = [&](int s) {
      if (s < nst) {
        const float4* g = reinterpret_cast<const float4*>(src + (size_t)s * kStageFloats);
        float4* d = reinterpret_cast<float4*>(sm + (size_t)(s % kStages) * kStageFloats);
        for (int x = tid; x < kStageFloats / 4; x += T) cp_async16(d + x, g + x);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) prefetch(s);
    for (int s = 0; s < nst; ++s) {
      cp_async_wait<kStages - 2>();
      __syncthreads();
      prefetch(s + kStages - 1);
      const float* __restrict__ q_s = sm + (size_t)(s % kStages) * kStageFloats;
      const int m = min(SJ, cnt - s * SJ);
      bool flag = false;
#pragma unroll U
      for (int j = 0; j < m; ++j) {
        float2 q[H], hh;
        ld_image<H>(q_s + j * Dq, q, hh);
#pragma unroll
        for (int r = 0; r < R; ++r) flag |= repel_dot<H, kSeries>(p[r], sp[r], thr[r], q, hh, acc[r], W[r]);
      }
      if (flag) {                                   // rare: the near pairs of this stage, in the difference form
        if (!fixed) {
#pragma unroll
          for (int x = 0; x < R * H; ++x) fix_s[x * T + tid] = make_float2(0.f, 0.f);
          fixed = true;
        }
        for (int j = 0; j < m; ++j) {
          float2 q[H], hh;
          ld_image<H>(q_s + j * Dq, q, hh);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float d2;
            if (near_test<H>(p[r], sp[r], thr[r], q, hh, d2)) {
              const float2 mh = make_float2(-0.5f, -0.5f);
              float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int k = 0; k < H; ++k) {
                const float2 dl = __ffma2_rn(q[k], mh, make_float2(-p[r][k].x, -p[r][k].y));   // (q - centre) - (p - centre), exact halving
                s2 = __ffma2_rn(dl, dl, s2);
              }
              const float ds = sqrt_approx(s2.x + s2.y) + 0.01f;
              const float w = ex2_approx(-3.0f * lg2_approx(ds));
              const float2 ww = make_float2(w, w);
#pragma unroll
              for (int k = 0; k < H; ++k) {
                const float2 dl = __ffma2_rn(q[k], mh, make_float2(-p[r][k].x, -p[r][k].y));
                float2* f = fix_s + (r * H + k) * T + tid;
                *f = __ffma2_rn(dl, ww, *f);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 o[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        // sum_j w (q_j - p) = -A / 2 - p W  (A accumulated on q~ = -2 (q - centre))
        o[k] = make_float2(fmaf(-p[r][k].x, W[r], -0.5f * acc[r][k].x), fmaf(-p[r][k].y, W[r], -0.5f * acc[r][k].y));
        if (fixed) o[k] = __fadd2_rn(o[k], fix_s[(r * H + k) * T + tid]);
      }
      st_point<H>(dv.rpart + ((size_t)c * dv.rows + lrow0 + tid + r * T) * Dp, o);
    }
    __syncthreads();   // the stages are refilled by the next item's prologue; item_s is rewritten
  }
}

#include "rowblock_tc.cuh"
#include "rowblock_tc2.cuh"

// ---------------------------------------------------------------------------------------------------
// the walk over the records of the own rows (springs and edge MAE)
// ---------------------------------------------------------------------------------------------------
// One thread per row, one warp per slice of 32 rows; step s of a warp handles record at(s) of each of its 32
// rows.  The partner rows are gathered into a per-warp shared-memory ring kRing steps ahead by cp.async, the
// four lanes of a quad fetching the 16-byte pieces of each other's rows: a warp-wide gather instruction then
// touches 8 cache lines instead of 32 (scattered 16-byte loads are bound by the L1 tag stage, one line per clock,
// long before L2 bandwidth), nothing is held in registers while in flight, and the records themselves ride the
// same ring (8 bytes per lane, one contiguous 256-byte line per warp and step).
constexpr int kRecAhead = 32;   // steps the record stream is prefetched into L2 ahead of its use
constexpr unsigned kSlotMask = 0x3fffffffu;

// kRing = steps of partner rows in flight per warp: 4 when the chip is full of warps (throughput: more CTAs per
// SM), 8 when a rank owns few rows (latency: a step then costs its dependent arithmetic, not an L2 round trip / 3).
template <int H, int kRing>
struct Ring {
  static constexpr int kRecRing = 2 * kRing;
  static constexpr int Dp = Row<H>::kStride;
  static constexpr int NC = Dp / 4;                         // 16-byte pieces of a row
  static constexpr int kRowBytes = NC * 16;
  static constexpr int kSlotBytes = 32 * kRowBytes;
  static constexpr int kWarpBytes = kRing * kSlotBytes + kRecRing * 32 * 8;
  // piece c of the row staged for `lane`.  Rows are stored lane-transposed ((lane & 3) * 8 + lane / 4): the 32
  // copies of one gather instruction (fixed quad member u, all quads, all pieces) then fill one contiguous
  // 8-row block - no bank conflict on the write side (the natural order gave 8-way conflicts: the L1 data pipe
  // was the limiter of the walk) - and 64-byte rows rotate their pieces by lane & 3 so that the eight lanes of a
  // 128-bit read phase hit eight different bank groups.
  static TL_D int piece(int lane, int c) {
    const int row = (lane & 3) * 8 + (lane >> 2);
    return NC == 4 ? row * 64 + ((c ^ (lane & 3)) << 4) : row * kRowBytes + (c << 4);
  }
};

TL_D void cp_async16_ca(unsigned smem_addr, const void* gmem) {
  // through L1: the two 16-byte halves of a 32-byte sector are asked for by two lanes of the same instruction;
  // the L1-bypassing form (.cg) fetches the sector once per lane (measured: twice the L2 traffic)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}

TL_D void keep(unsigned& x) { asm volatile("" : "+r"(x)); }   // opaque to the compiler: computed once, not rematerialised in the loop
TL_D uint4 lds128(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
TL_D uint2 lds64(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }

// consume(record, partner row) is called for every step in order, by all lanes of the warp together.
// Order inside a step (a lone warp per scheduler issues in order, so the order is the schedule): wait for the
// step's data, read it (row pieces, own record, the four records of the quad for the gather) in one burst of
// shared-memory loads, issue the gathers of step t + kRing - 1 and the record of step t + 2 kRing - 1, then the
// arithmetic of step t, which the compiler is free to spread between the copy instructions.
template <int H, int kRing, class F>
TL_D void walk_rows(const float* __restrict__ P, const uint2* __restrict__ rec /* + lane */, int width, int start, char* wsm,
                    F&& consume) {
  typedef Ring<H, kRing> R;
  constexpr int kRecRing = R::kRecRing;
  constexpr unsigned kRowsBytes = kRing * R::kSlotBytes, kRecSlotBytes = 32 * 8;
  const int lane = threadIdx.x & 31;
  const unsigned rows_a = (unsigned)__cvta_generic_to_shared(wsm);
  const unsigned recs_a = rows_a + kRowsBytes;
  const int c = lane & 3, quad = lane & ~3;
  // lane constants (byte offsets inside a ring slot)
  unsigned dst_off[4], own_off[R::NC];
#pragma unroll
  for (int u = 0; u < 4; ++u) { dst_off[u] = (unsigned)R::piece(quad + u, c); keep(dst_off[u]); }
#pragma unroll
  for (int k = 0; k < R::NC; ++k) { own_off[k] = (unsigned)R::piece(lane, k); keep(own_off[k]); }
  unsigned rec_own = (unsigned)lane * 8u, rec_quad = (unsigned)quad * 8u;
  keep(rec_own); keep(rec_quad);
  const char* P_c = reinterpret_cast<const char*>(P) + c * 16;
  asm volatile("" : "+l"(P_c));
  asm volatile("" : "+l"(rec));
  // cursors (warp-uniform)
  int r_at = start, r_left = width;
  unsigned r_slot_a = recs_a;                                  // record-ring slot the next record goes to
  int pf_at = start + kRecAhead;                               // record line pulled into L2 this step (lanes 0, 1)
  if (width > 0) pf_at %= width;
  const uint2* rec_line = rec - lane + (lane & 1) * 16;
  auto issue_rec = [&]() {
    if (r_left > 0) {
      if (lane < 2 && r_left > kRecAhead) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec_line + (size_t)pf_at * 32));
      pf_at = (pf_at + 1 == width) ? 0 : pf_at + 1;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(r_slot_a + rec_own), "l"(rec + (size_t)r_at * 32) : "memory");
      r_at = (r_at + 1 == width) ? 0 : r_at + 1;
      --r_left;
    }
    r_slot_a = (r_slot_a + kRecSlotBytes == recs_a + kRecRing * kRecSlotBytes) ? recs_a : r_slot_a + kRecSlotBytes;
  };
  int g_left = width;
  unsigned g_slot_a = rows_a, g_rslot_a = recs_a;               // row-ring slot gathered into, record slot it reads
  auto issue_gather = [&](const uint4 ja, const uint4 jb) {     // the four records of the quad: (ja.x, ja.z, jb.x, jb.z) are the .x words
    if (g_left > 0) {
      if (c < R::NC) {
        const unsigned j[4] = {ja.x & kSlotMask, ja.z & kSlotMask, jb.x & kSlotMask, jb.z & kSlotMask};
#pragma unroll
        for (int u = 0; u < 4; ++u) cp_async16_ca(g_slot_a + dst_off[u], P_c + (size_t)j[u] * (R::Dp * 4));
      }
      --g_left;
    }
    g_slot_a = (g_slot_a + R::kSlotBytes == rows_a + kRowsBytes) ? rows_a : g_slot_a + R::kSlotBytes;
    g_rslot_a = (g_rslot_a + kRecSlotBytes == recs_a + kRecRing * kRecSlotBytes) ? recs_a : g_rslot_a + kRecSlotBytes;
  };
  for (int st = 0; st < kRecRing - 1; ++st) issue_rec();
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
  for (int st = 0; st < kRing - 1; ++st) {
    const uint4 ja = lds128(g_rslot_a + rec_quad), jb = lds128(g_rslot_a + rec_quad + 16);
    issue_gather(ja, jb);
    cp_async_commit();
  }
  unsigned slot_a = rows_a, rslot_a = recs_a;
  for (int t = 0; t < width; ++t) {
    cp_async_wait<kRing - 2>();
    __syncwarp();
    // ---- everything this step reads from shared memory ----
    uint4 qv[R::NC];
#pragma unroll
    for (int k = 0; k < H / 2; ++k) qv[k] = lds128(slot_a + own_off[k]);
    uint2 qt = make_uint2(0u, 0u);
    if constexpr (H % 2 == 1) qt = lds64(slot_a + own_off[H / 2]);
    const uint2 r = lds64(rslot_a + rec_own);
    const uint4 ja = lds128(g_rslot_a + rec_quad), jb = lds128(g_rslot_a + rec_quad + 16);
    // ---- copies for the steps ahead ----
    issue_gather(ja, jb);
    issue_rec();
    cp_async_commit();
    slot_a = (slot_a + R::kSlotBytes == rows_a + kRowsBytes) ? rows_a : slot_a + R::kSlotBytes;
    rslot_a = (rslot_a + kRecSlotBytes == recs_a + kRecRing * kRecSlotBytes) ? recs_a : rslot_a + kRecSlotBytes;
    // ---- this step ----
    float2 q[H];
#pragma unroll
    for (int k = 0; k < H / 2; ++k) {
      q[2 * k] = make_float2(__uint_as_float(qv[k].x), __uint_as_float(qv[k].y));
      q[2 * k + 1] = make_float2(__uint_as_float(qv[k].z), __uint_as_float(qv[k].w));
    }
    if constexpr (H % 2 == 1) q[H - 1] = make_float2(__uint_as_float(qt.x), __uint_as_float(qt.y));
    consume(r, q);
  }
  cp_async_wait<0>();
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// springs: one thread per own row, Gauss-Seidel along the row
// ---------------------------------------------------------------------------------------------------
template <int H, int kRing>
__global__ void __launch_bounds__(kBlockRows) spring_kernel(RowDev dv, FitParams prm, int cur, unsigned epoch) {
  constexpr int Dp = Row<H>::kStride;
  extern __shared__ float4 smem4[];
  asm volatile("griddepcontrol.launch_dependents;");   // a repulsion pass launched as programmatic dependent may start now
  if (__ldcg(&dv.state->stop)) return;
  if (!wait_epoch(dv, epoch)) { if (blockIdx.x == 0 && threadIdx.x == 0) peer_timeout(dv); return; }
  const int tid = threadIdx.x;
  const float* __restrict__ P = dv.pos[dv.rank] + (size_t)cur * dv.cap_rows * Dp;
  const float k = (float)__ldcg(&dv.state->k);
  const int iter = __ldcg(&dv.state->iter);
  for (int blk = blockIdx.x; blk < dv.rows / kBlockRows; blk += gridDim.x) {
  const int lrow = blk * kBlockRows + tid;                 // row inside the own block
  const int row = dv.row0 + lrow;                          // slot
  const float dp1 = dv.dp1[row];                           // 0 for a padding row: its lane walks along, stores nothing
  const float safe = dp1 > 0.f ? dp1 : 1.f;
  float2 p0n[H], xn[H];   // -P_i and -x (the running position, negated: deltas are q + (-x))
  {
    float2 p[H];
    ld_point<H>(P + (size_t)row * Dp, p);
#pragma unroll
    for (int kk = 0; kk < H; ++kk) { p0n[kk] = make_float2(-p[kk].x, -p[kk].y); xn[kk] = p0n[kk]; }
  }
  const float rdeg = (float)(0.5 * prm.c_repulsion) / safe;
  const float two_k_rnorm = 2.0f * k / (4.0f * safe + k);
  const int slice = lrow >> 5;
  const int width = dv.swidth[slice];
  int start = 0;
  if (width > 0) start = (int)(mix64(dv.seed ^ mix64(((unsigned long long)(unsigned)iter << 32) | (unsigned)(row >> 5))) % (unsigned long long)width);
  walk_rows<H, kRing>(P, dv.recs + dv.soff[slice] + (tid & 31), width, start,
               reinterpret_cast<char*>(smem4) + (size_t)(tid >> 5) * Ring<H, kRing>::kWarpBytes, [&](const uint2 r, const float2 (&q)[H]) {
    const unsigned type = r.x >> 30;
    const float target = __uint_as_float(r.y);
    float2 dl[H], d0[H];
    const float d2 = dist2<H>(xn, q, dl);
    const float e2 = dist2<H>(p0n, q, d0);
    const float dist = sqrt_approx(d2);
    const bool spring = type == 0u || (type == 1u ? dist < target : (type == 2u && dist > target));   // :237-243
    const float ds0 = sqrt_approx(e2) + 0.01f;
    const float w0 = spring ? rdeg * rcp_approx(ds0 * ds0 * ds0) : 0.f;
    const float f = spring ? two_k_rnorm * (target - dist) * rcp_approx(dist + 0.01f) : 0.f;
    // x += -delta f + d0 w0  <=>  (-x) += delta f - d0 w0; the take-back term does not wait for f
    const float2 ff = make_float2(f, f), nw = make_float2(-w0, -w0);
#pragma unroll
    for (int kk = 0; kk < H; ++kk) xn[kk] = __ffma2_rn(dl[kk], ff, __ffma2_rn(d0[kk], nw, xn[kk]));
  });
  if (dp1 > 0.f) {
    float2 x[H];
#pragma unroll
    for (int kk = 0; kk < H; ++kk) x[kk] = make_float2(-xn[kk].x, -xn[kk].y);
    st_point<H>(dv.xs + (size_t)lrow * Dp, x);
  }
  }
}

// ---------------------------------------------------------------------------------------------------
// combine: P'_i = (position after the springs) + (repulsion of the iteration), stored into every replica
// ---------------------------------------------------------------------------------------------------
// The spring walk (a long dependent chain per row: latency) and the repulsion pass (FMA throughput) both read
// only the iteration's snapshot, so they run side by side (the repulsion pass is a programmatic dependent launch of
// the walk: it starts as soon as the walk's CTAs are resident); this kernel, a normal launch, joins them.  The
// repulsion sums of a row are added chunk by chunk in a fixed order (the result does not depend on the rank count).
template <int H>
__global__ void __launch_bounds__(kBlockRows) combine_kernel(RowDev dv, FitParams prm, int cur, unsigned epoch) {
  constexpr int Dp = Row<H>::kStride;
  if (__ldcg(&dv.state->stop)) return;
  const int tid = threadIdx.x;
  const int lrow = blockIdx.x * kBlockRows + tid;
  const int row = dv.row0 + lrow;
  const float dp1 = dv.dp1[row];
  if (dp1 > 0.f) {
    float2 x[H], rs[H];
    ld_point<H>(dv.xs + (size_t)lrow * Dp, x);
#pragma unroll
    for (int kk = 0; kk < H; ++kk) rs[kk] = make_float2(0.f, 0.f);
    const int nparts = (dv.adaptive && !form_is_tensor(dv)) ? dv.chunks : dv.rparts;   // the FP32 form writes one entry per chunk
    for (int c = 0; c < nparts; ++c) {
      float2 t[H];
      ld_point<H>(dv.rpart + ((size_t)c * dv.rows + lrow) * Dp, t);
#pragma unroll
      for (int kk = 0; kk < H; ++kk) rs[kk] = __fadd2_rn(rs[kk], t[kk]);
    }
    const float nrdeg = -(float)(0.5 * prm.c_repulsion) / dp1;     // R_i = -sum * (c / 2) / (deg_i + 1)
    const float2 rr = make_float2(nrdeg, nrdeg);
#pragma unroll
    for (int kk = 0; kk < H; ++kk) x[kk] = __ffma2_rn(rs[kk], rr, x[kk]);
    const size_t o = ((size_t)(cur ^ 1) * dv.cap_rows + row) * Dp;
    for (int g = 0; g < dv.G; ++g) st_point<H>(dv.pos[g] + o, x);   // own replica and every peer (NVLink stores)
  }
  if (last_cta(&dv.counters[0])) {
    if (tid == 0) {
      FitState* st = dv.state;
      const int iter = st->iter;
      st->k = st->k * (1.0 - prm.cooling_rate);         // :289
      st->iter = iter + 1;
      st->pair_updates += dv.pairs_per_iter;
      dv.counters[2] = 0u;                              // the repulsion pass of the next iteration starts at item 0
      if (dv.host_flag) dv.host_flag[1] = iter + 1;
      __threadfence_system();
      signal_epoch(dv, epoch);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// edge MAE on the new positions: one thread per own row over the records this rank counts
// ---------------------------------------------------------------------------------------------------
template <int H, int kRing>
__global__ void __launch_bounds__(kBlockRows) mae_kernel(RowDev dv, int nxt, unsigned wait_for, unsigned epoch) {
  constexpr int Dp = Row<H>::kStride;
  extern __shared__ float4 smem4[];
  __shared__ double red_s[3][kBlockRows / 32];
  if (__ldcg(&dv.state->stop)) return;
  if (!wait_epoch(dv, wait_for)) { if (blockIdx.x == 0 && threadIdx.x == 0) peer_timeout(dv); return; }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lrow = blockIdx.x * kBlockRows + tid;
  const int row = dv.row0 + lrow;
  const float* __restrict__ P = dv.pos[dv.rank] + (size_t)nxt * dv.cap_rows * Dp;
  double sum = 0.0, cnt = 0.0, bad = 0.0;
  float2 pn[H];
  {
    float2 p[H];
    ld_point<H>(P + (size_t)row * Dp, p);
#pragma unroll
    for (int kk = 0; kk < H; ++kk) {
      if (!isfinite(p[kk].x) || !isfinite(p[kk].y)) bad = 1.0;
      pn[kk] = make_float2(-p[kk].x, -p[kk].y);
    }
  }
  const int slice = lrow >> 5;
  walk_rows<H, kRing>(P, dv.mrecs + dv.moff[slice] + lane, dv.mwidth[slice], 0,
               reinterpret_cast<char*>(smem4) + (size_t)warp * Ring<H, kRing>::kWarpBytes, [&](const uint2 r, const float2 (&q)[H]) {
    float2 dl[H];
    const float dist = __fsqrt_rn(dist2<H>(pn, q, dl));
    const unsigned type = r.x >> 30;
    const float target = __uint_as_float(r.y);
    const bool on = type == 0u || (type == 1u ? dist < target : (type == 2u && dist > target));   // :72-75
    if (on) { sum += fabs((double)target - (double)dist); cnt += 1.0; }
  });
  // fixed tree: lanes, then warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    bad += __shfl_down_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) { red_s[0][warp] = sum; red_s[1][warp] = cnt; red_s[2][warp] = bad; }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, c = 0.0, b = 0.0;
    for (int w = 0; w < kBlockRows / 32; ++w) { s += red_s[0][w]; c += red_s[1][w]; b += red_s[2][w]; }
    const size_t e = (size_t)(row / kBlockRows) * 4;
    for (int g = 0; g < dv.G; ++g) {
      double* o = dv.red[g] + e;
      o[0] = s; o[1] = c; o[2] = b;
    }
  }
  if (last_cta(&dv.counters[1]) && tid == 0) signal_epoch(dv, epoch);
}

// ---------------------------------------------------------------------------------------------------
// controller (one CTA): pooled MAE, three-way classification, finite check
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctl_kernel(RowDev dv, FitParams prm, int check, int fin, unsigned wait_for) {
  __shared__ double red_s[3][256];
  if (__ldcg(&dv.state->stop)) return;
  if (!wait_epoch(dv, wait_for)) { if (threadIdx.x == 0) peer_timeout(dv); return; }
  const int tid = threadIdx.x;
  const int entries = dv.slots / kBlockRows;
  const double* __restrict__ red = dv.red[dv.rank];
  double s = 0.0, c = 0.0, b = 0.0;
  for (int e = tid; e < entries; e += 256) { s += __ldcg(red + (size_t)e * 4); c += __ldcg(red + (size_t)e * 4 + 1); b += __ldcg(red + (size_t)e * 4 + 2); }
  red_s[0][tid] = s; red_s[1][tid] = c; red_s[2][tid] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { red_s[0][tid] += red_s[0][tid + o]; red_s[1][tid] += red_s[1][tid + o]; red_s[2][tid] += red_s[2][tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    FitState st = *dv.state;
    const int iter = st.iter - 1;   // the iteration that just finished
    st.snapshot = 0;
    if (check) {
      controller_check(st, prm, iter, red_s[0][0], (long long)red_s[1][0]);
      if (dv.trace && iter >= 0 && iter < prm.n_iter) dv.trace[iter] = st.last_error;
    }
    if (!st.stop && fin && red_s[2][0] > 0.0) { st.status = 2; st.fail_iter = iter + 1; st.stop = 1; }
    *dv.state = st;
    if (dv.host_flag && st.stop) { __threadfence_system(); dv.host_flag[0] = 1; }
  }
}

__global__ void __launch_bounds__(256) snap_kernel(RowDev dv, int nxt) {
  if (!__ldcg(&dv.state->snapshot)) return;
  const size_t total4 = (size_t)dv.slots * dv.Dp / 4;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(dv.pos[dv.rank] + (size_t)nxt * dv.cap_rows * dv.Dp);
  float4* dst = reinterpret_cast<float4*>(dv.best);
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total4; x += (size_t)gridDim.x * blockDim.x) dst[x] = src[x];
}

// ---------------------------------------------------------------------------------------------------
// set-up kernels: COO edge list -> the two SELL-32 record arrays of the own rows
// ---------------------------------------------------------------------------------------------------
// A pair is counted in the MAE by exactly one of its endpoints, the two halves of every row about equal.
TL_HD bool counts_here(uint32_t self, uint32_t other) { return (((self + other) & 1u) != 0u) == (self < other); }

__global__ void count_kernel(const int32_t* __restrict__ ei, const int32_t* __restrict__ ej, long long E, long long n,
                             const int32_t* __restrict__ slot_of_point, int row0, int rows, unsigned* len, unsigned* mlen, int* bad) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long a = ei[e], b = ej[e];
    if (a < 0 || b < 0 || a >= n || b >= n || a == b) { *bad = 1; continue; }
    const uint32_t sa = (uint32_t)slot_of_point[a], sb = (uint32_t)slot_of_point[b];
    const int la = (int)sa - row0, lb = (int)sb - row0;
    if (la >= 0 && la < rows) { atomicAdd(&len[la], 1u); if (counts_here(sa, sb)) atomicAdd(&mlen[la], 1u); }
    if (lb >= 0 && lb < rows) { atomicAdd(&len[lb], 1u); if (counts_here(sb, sa)) atomicAdd(&mlen[lb], 1u); }
  }
}
__global__ void fill_kernel(const int32_t* __restrict__ ei, const int32_t* __restrict__ ej, const double* __restrict__ dist,
                            const int32_t* __restrict__ thr, long long E, const int32_t* __restrict__ slot_of_point, int row0,
                            int rows, const unsigned long long* __restrict__ off, unsigned* cursor, uint2* tmp) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const uint32_t sa = (uint32_t)slot_of_point[ei[e]], sb = (uint32_t)slot_of_point[ej[e]];
    const int t = thr[e];
    const unsigned type = t == 0 ? 0u : (t == 1 ? 1u : 2u);   // src/optimization.cpp:237-243
    const unsigned tb = __float_as_uint((float)dist[e]);
    const int la = (int)sa - row0, lb = (int)sb - row0;
    if (la >= 0 && la < rows) tmp[off[la] + atomicAdd(&cursor[la], 1u)] = make_uint2(sb | (type << 30), tb);
    if (lb >= 0 && lb < rows) tmp[off[lb] + atomicAdd(&cursor[lb], 1u)] = make_uint2(sa | (type << 30), tb);
  }
}
// One warp per own row: sort the row's records by (partner, type, target bits) - a fixed order whatever the
// scatter left - and write them, transposed, into the slices; the tail of a row up to the slice width is
// padding (type 3: never a spring, never counted; partner = the row itself).  Rows of up to kSortCap records
// are sorted in shared memory (bitonic network on the 64-bit keys, from which the record is rebuilt), longer
// ones by counting ranks in global memory.
constexpr int kSortCap = 2048;
TL_D unsigned long long rec_key(uint2 r) {
  return ((unsigned long long)(r.x & 0x3fffffffu) << 34) | ((unsigned long long)(r.x >> 30) << 32) | r.y;
}
TL_D uint2 key_rec(unsigned long long k) {
  return make_uint2((unsigned)(k >> 34) | ((unsigned)((k >> 32) & 3ull) << 30), (unsigned)k);
}
__global__ void __launch_bounds__(128) sell_kernel(const uint2* __restrict__ tmp, const unsigned long long* __restrict__ off,
                                                   const unsigned* __restrict__ len, int row0, int rows,
                                                   const unsigned long long* __restrict__ soff, const int* __restrict__ swidth,
                                                   const unsigned long long* __restrict__ moff, const int* __restrict__ mwidth,
                                                   uint2* recs, uint2* mrecs) {
  extern __shared__ unsigned long long sort_s[];   // [4][kSortCap]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lrow = blockIdx.x * 4 + warp;
  if (lrow >= rows) return;
  const unsigned n = len[lrow];
  const uint2* __restrict__ in = tmp + off[lrow];
  const uint32_t self = (uint32_t)(row0 + lrow);
  const int slice = lrow >> 5, rl = lrow & 31;
  uint2* out = recs + soff[slice] + rl;
  uint2* mout = mrecs + moff[slice] + rl;
  unsigned mcount = 0;
  if (n <= (unsigned)kSortCap) {
    unsigned long long* ks = sort_s + (size_t)warp * kSortCap;
    unsigned m = 32;
    while (m < n) m <<= 1;
    for (unsigned x = lane; x < m; x += 32) ks[x] = x < n ? rec_key(in[x]) : ~0ull;
    __syncwarp();
    for (unsigned k = 2; k <= m; k <<= 1)
      for (unsigned j = k >> 1; j > 0; j >>= 1) {
        for (unsigned i = lane; i < m; i += 32) {
          const unsigned l = i ^ j;
          if (l > i) {
            const unsigned long long a = ks[i], b = ks[l];
            if ((a > b) == ((i & k) == 0)) { ks[i] = b; ks[l] = a; }
          }
        }
        __syncwarp();
      }
    for (unsigned x0 = 0; x0 < n; x0 += 32) {
      const unsigned x = x0 + lane;
      const uint2 r = x < n ? key_rec(ks[x]) : make_uint2(0u, 0u);
      const bool counted = x < n && counts_here(self, r.x & 0x3fffffffu);
      const unsigned vote = __ballot_sync(0xffffffffu, counted);
      if (x < n) out[(size_t)x * 32] = r;
      if (counted) mout[(size_t)(mcount + __popc(vote & ((1u << lane) - 1u))) * 32] = r;
      mcount += __popc(vote);
    }
  } else {
    for (unsigned x0 = 0; x0 < n; x0 += 32) {
      const unsigned x = x0 + lane;
      if (x < n) {
        const uint2 r = in[x];
        const unsigned long long mkey = rec_key(r);
        const bool mcounted = counts_here(self, r.x & 0x3fffffffu);
        unsigned rank = 0, mrank = 0;
        for (unsigned y = 0; y < n; ++y) {
          const uint2 o = in[y];
          const unsigned long long okey = rec_key(o);
          const bool before = okey < mkey || (okey == mkey && y < x);
          rank += before ? 1u : 0u;
          mrank += (before && counts_here(self, o.x & 0x3fffffffu)) ? 1u : 0u;
        }
        out[(size_t)rank * 32] = r;
        if (mcounted) mout[(size_t)mrank * 32] = r;
      }
    }
    for (unsigned y = lane; y < n; y += 32) mcount += counts_here(self, in[y].x & 0x3fffffffu) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mcount += __shfl_xor_sync(0xffffffffu, mcount, o);
  }
  const uint2 pad = make_uint2(self | (3u << 30), 0u);
  for (unsigned x = n + lane; x < (unsigned)swidth[slice]; x += 32) out[(size_t)x * 32] = pad;
  for (unsigned x = mcount + lane; x < (unsigned)mwidth[slice]; x += 32) mout[(size_t)x * 32] = pad;
}

template <class F>
void dispatch_h(int H, F&& f) {
  switch (H) {
    case 1: f(std::integral_constant<int, 1>()); break;
    case 2: f(std::integral_constant<int, 2>()); break;
    case 3: f(std::integral_constant<int, 3>()); break;
    case 4: f(std::integral_constant<int, 4>()); break;
    case 5: f(std::integral_constant<int, 5>()); break;
    case 6: f(std::integral_constant<int, 6>()); break;
    case 7: f(std::integral_constant<int, 7>()); break;
    case 8: f(std::integral_constant<int, 8>()); break;
    default: throw BadArg("row-block mode supports ndim <= 16");
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct RowHandle {            // what a rank publishes: its shared block and the geometry the others must agree with
  cudaIpcMemHandle_t mem;
  int32_t rank, n_ranks, slots, Dp;
  uint64_t layout;            // byte size of the shared block
};

struct RowPlan {
  int device = 0;
  int64_t n = 0, E = 0;
  RowDev dv{};
  FitParams prm{};
  std::vector<int32_t> slot_of_point;
  // shared block (cudaMalloc: exportable): positions [2][cap_rows][Dp] | red [slots/128][4] | flags
  char* shared = nullptr;
  size_t shared_bytes = 0, off_red = 0, off_flags = 0;
  void* peer_base[kMaxShards] = {};   // mapped blocks of the other ranks (IPC) - closed on destroy
  bool peer_ipc[kMaxShards] = {};
  // local
  float* best = nullptr; float* dp1 = nullptr; float* rpart = nullptr; float* xs = nullptr;
  float* img = nullptr; float* hmax = nullptr;
  float* tc_buf = nullptr;    // the five arrays of TcImage in one allocation
  bool tc_form = false;       // distances on the tensor cores (image_tc_kernel + repulse_tc_kernel)
  bool tc2_form = false;      // distances and accumulation on the tensor cores (+ image_t2_kernel, repulse_tc2_kernel)
  int tc_cpi = 1;             // partner chunks per work item
  bool dot_form = false;      // repulsion in the inner-product form (image_kernel + repulse_dot_kernel)
  uint2* recs = nullptr; uint2* mrecs = nullptr;
  unsigned long long* soff = nullptr; unsigned long long* moff = nullptr; int* swidth = nullptr; int* mwidth = nullptr;
  FitState* state = nullptr; double* trace = nullptr; unsigned* counters = nullptr;
  volatile int* h_flag = nullptr; int* d_flag = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t n_holdout = 0;
  std::vector<int32_t> hold_i, hold_j; std::vector<double> hold_truth;
  int iter_launched = 0;      // iterations enqueued so far
  unsigned epoch = 0;         // signals enqueued so far (the same number on every rank)
  int64_t launches = 0;
  int64_t n_recs = 0, n_mrecs = 0;
  int64_t tensor_iters = -1;  // adaptive runs: iterations that ran the tensor form (read back by row_result; -1 = not adaptive / not read yet)
  int rep_ctas = 0, rep_threads = 128;
  double near_limit = 0.0;    // adaptive runs: the tensor form while the probed near-pair fraction is at most this
  int rep_form = 0;           // shape of the repulsion pass in use (repulse_variant / configure_repulse)
  bool deep_ring = false;
  bool overlap = true;        // repulsion pass launched as programmatic dependent of the spring walk (TOPOLOW_OVERLAP=0: in order)
  int sm_count = 148;
  void* rep_fn = nullptr;
  size_t rep_smem = 0;
  int f32_ctas = 0, f32_threads = 128;   // the FP32 difference form beside a tensor form (adaptive runs launch both)
  size_t f32_smem = 0;
  double total_ms = 0.0;
  bool attached = false;

  ~RowPlan() {
    DeviceScope on(device);
    if (stream) cudaStreamSynchronize(stream);
    for (int q = 0; q < kMaxShards; ++q) if (peer_base[q] && peer_ipc[q]) cudaIpcCloseMemHandle(peer_base[q]);
    if (shared) cudaFree(shared);
    pool_free(best); pool_free(dp1); pool_free(rpart); pool_free(xs); pool_free(img); pool_free(hmax); pool_free(tc_buf); pool_free(recs); pool_free(mrecs);
    pool_free(soff); pool_free(moff); pool_free(swidth); pool_free(mwidth);
    pool_free(state); pool_free(trace); pool_free(counters);
    if (h_flag) cudaFreeHost((void*)h_flag);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void set_self_pointers(RowPlan& rp) {
  RowDev& dv = rp.dv;
  dv.pos[dv.rank] = reinterpret_cast<float*>(rp.shared);
  dv.red[dv.rank] = reinterpret_cast<double*>(rp.shared + rp.off_red);
  dv.flags[dv.rank] = reinterpret_cast<unsigned*>(rp.shared + rp.off_flags);
}
void set_peer_pointers(RowPlan& rp, int q, char* base) {
  rp.dv.pos[q] = reinterpret_cast<float*>(base);
  rp.dv.red[q] = reinterpret_cast<double*>(base + rp.off_red);
  rp.dv.flags[q] = reinterpret_cast<unsigned*>(base + rp.off_flags);
}

// Builds the record arrays of the own rows on the device.
void build_records(RowPlan& rp, const topolow_problem& pb) {
  RowDev& dv = rp.dv;
  cudaStream_t s = rp.stream;
  const long long E = pb.n_edges;
  const int rows = dv.rows, slices = rows / 32;
  AsyncBuf<int32_t> d_ei(E, s), d_ej(E, s), d_thr(E, s), d_slot(rp.slot_of_point.size(), s);
  AsyncBuf<double> d_dist(E, s);
  AsyncBuf<unsigned> d_len(rows, s), d_mlen(rows, s), d_cur(rows, s);
  AsyncBuf<unsigned long long> d_off(rows + 1, s);
  AsyncBuf<int> d_bad(1, s);
  PhaseTimer pt(s);
  TL_CUDA(cudaMemcpyAsync(d_slot, rp.slot_of_point.data(), rp.slot_of_point.size() * 4, cudaMemcpyHostToDevice, s));
  if (E > 0) {
    TL_CUDA(cudaMemcpyAsync(d_ei, pb.edge_i, E * 4, cudaMemcpyHostToDevice, s));
    TL_CUDA(cudaMemcpyAsync(d_ej, pb.edge_j, E * 4, cudaMemcpyHostToDevice, s));
    TL_CUDA(cudaMemcpyAsync(d_dist, pb.edge_dist, E * 8, cudaMemcpyHostToDevice, s));
    TL_CUDA(cudaMemcpyAsync(d_thr, pb.edge_thresh, E * 4, cudaMemcpyHostToDevice, s));
  }
  TL_CUDA(cudaMemsetAsync(d_len, 0, rows * 4, s));
  TL_CUDA(cudaMemsetAsync(d_mlen, 0, rows * 4, s));
  TL_CUDA(cudaMemsetAsync(d_cur, 0, rows * 4, s));
  TL_CUDA(cudaMemsetAsync(d_bad, 0, 4, s));
  pt.mark("rows: host to device");
  const int blocks = 148 * 8, threads = 256;
  if (E > 0) {
    count_kernel<<<blocks, threads, 0, s>>>(d_ei, d_ej, E, pb.n, d_slot, dv.row0, rows, d_len, d_mlen, d_bad);
    TL_CUDA(cudaGetLastError());
  }
  std::vector<unsigned> len(rows), mlen(rows);
  int bad = 0;
  TL_CUDA(cudaMemcpyAsync(len.data(), d_len, rows * 4, cudaMemcpyDeviceToHost, s));
  TL_CUDA(cudaMemcpyAsync(mlen.data(), d_mlen, rows * 4, cudaMemcpyDeviceToHost, s));
  TL_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
  TL_CUDA(cudaStreamSynchronize(s));
  if (bad) throw BadArg("edge index out of range");
  std::vector<unsigned long long> off(rows + 1, 0), soff(slices), moff(slices);
  std::vector<int> swidth(slices), mwidth(slices);
  for (int r = 0; r < rows; ++r) off[r + 1] = off[r] + len[r];
  unsigned long long so = 0, mo = 0;
  for (int sl = 0; sl < slices; ++sl) {
    unsigned w = 0, mw = 0;
    for (int r = 0; r < 32; ++r) { w = std::max(w, len[sl * 32 + r]); mw = std::max(mw, mlen[sl * 32 + r]); }
    soff[sl] = so; moff[sl] = mo; swidth[sl] = (int)w; mwidth[sl] = (int)mw;
    so += (unsigned long long)w * 32; mo += (unsigned long long)mw * 32;
  }
  rp.n_recs = (int64_t)so; rp.n_mrecs = (int64_t)mo;
  pool_alloc(rp.recs, (size_t)so * sizeof(uint2));
  pool_alloc(rp.mrecs, (size_t)mo * sizeof(uint2));
  pool_alloc(rp.soff, slices * sizeof(unsigned long long));
  pool_alloc(rp.moff, slices * sizeof(unsigned long long));
  pool_alloc(rp.swidth, slices * sizeof(int));
  pool_alloc(rp.mwidth, slices * sizeof(int));
  pool_ready();
  AsyncBuf<uint2> d_tmp((size_t)off[rows], s);
  TL_CUDA(cudaMemcpyAsync(d_off, off.data(), (rows + 1) * 8, cudaMemcpyHostToDevice, s));
  TL_CUDA(cudaMemcpyAsync(rp.soff, soff.data(), slices * 8, cudaMemcpyHostToDevice, s));
  TL_CUDA(cudaMemcpyAsync(rp.moff, moff.data(), slices * 8, cudaMemcpyHostToDevice, s));
  TL_CUDA(cudaMemcpyAsync(rp.swidth, swidth.data(), slices * 4, cudaMemcpyHostToDevice, s));
  TL_CUDA(cudaMemcpyAsync(rp.mwidth, mwidth.data(), slices * 4, cudaMemcpyHostToDevice, s));
  pt.mark("rows: lengths + offsets");
  if (E > 0) {
    fill_kernel<<<blocks, threads, 0, s>>>(d_ei, d_ej, d_dist, d_thr, E, d_slot, dv.row0, rows, d_off, d_cur, d_tmp);
    TL_CUDA(cudaGetLastError());
  }
  // per device, so every time (a function-local static would configure only the device of the first call)
  TL_CUDA(cudaFuncSetAttribute(sell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSortCap * 8));
  sell_kernel<<<(rows + 3) / 4, 128, 4 * kSortCap * 8, s>>>(d_tmp, d_off, d_len, dv.row0, rows, rp.soff, rp.swidth, rp.moff, rp.mwidth,
                                              rp.recs, rp.mrecs);
  TL_CUDA(cudaGetLastError());
  TL_CUDA(cudaStreamSynchronize(s));   // the host vectors above go out of scope
  pt.mark("rows: fill + sort + transpose");
  dv.recs = rp.recs; dv.mrecs = rp.mrecs; dv.soff = rp.soff; dv.moff = rp.moff; dv.swidth = rp.swidth; dv.mwidth = rp.mwidth;
}

TcImage tc_image(const RowPlan& rp) {
  TcImage im;
  const size_t cap = rp.dv.cap_rows;
  im.xhi = rp.tc_buf; im.xlo = im.xhi + 16 * cap; im.aug_a = im.xlo + 16 * cap; im.aug_b = im.aug_a + 4 * cap; im.rows32 = im.aug_b + 4 * cap;
  return im;
}
T2Image t2_image(const RowPlan& rp) {
  T2Image im;
  im.a = tc_image(rp);
  im.yhi = im.a.rows32 + 16 * rp.dv.cap_rows; im.ylo = im.yhi + 32 * rp.dv.cap_rows;
  return im;
}
template <int H>
void launch_image_tc(RowPlan& rp, cudaStream_t s, int cur, unsigned e_prev) {
  image_tc_kernel<H><<<(unsigned)(rp.dv.cap_rows / kBlockRows), kBlockRows, 0, s>>>(rp.dv, tc_image(rp), cur, e_prev);
  rp.launches += 1;
  if (rp.tc2_form) {
    image_t2_kernel<<<(unsigned)(rp.dv.cap_rows / 4 * 32 / 256), 256, 0, s>>>(rp.dv, t2_image(rp));
    rp.launches += 1;
  }
}
template <int H>
void launch_repulse_tc(RowPlan& rp, cudaStream_t s, int cur, bool dependent) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)rp.rep_ctas); cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = TcSmem::kTotal; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = dependent ? 1 : 0;
  if (rp.tc2_form) {
    cfg.blockDim = dim3(kT2Threads); cfg.dynamicSmemBytes = T2Smem::kTotal;
    TL_CUDA(cudaLaunchKernelEx(&cfg, repulse_tc2_kernel<H>, rp.dv, t2_image(rp), cur, rp.tc_cpi));
    return;
  }
  TL_CUDA(cudaLaunchKernelEx(&cfg, repulse_tc_kernel<H>, rp.dv, tc_image(rp), cur, rp.tc_cpi));
}

template <int H, int kRing> constexpr size_t kWalkSmem = (size_t)(kBlockRows / 32) * Ring<H, kRing>::kWarpBytes;

// the spring / MAE launches of one rank (deep ring when the rank owns few rows)
template <int H>
void launch_spring(const RowPlan& rp, cudaStream_t s, int cur, unsigned e_spring);
template <int H>
void launch_mae(const RowPlan& rp, cudaStream_t s, int nxt, unsigned e_spring, unsigned e_mae);

// One iteration of one rank.  overlap: the repulsion pass starts while the spring walk is still running (ranks
// with few rows: the walk is a latency chain on a few warps); otherwise the kernels run one after another.
template <int H>
void launch_iteration(RowPlan& rp, cudaStream_t s, cudaEvent_t* ev /* 7 events or null */, bool overlap) {
  RowDev& dv = rp.dv;
  const int t = rp.iter_launched;
  const int cur = t & 1, nxt = cur ^ 1;
  const bool check = is_check_iter(t, rp.prm), fin = ((t + 1) % 10 == 0);
  const unsigned e_prev = rp.epoch;          // the epoch every rank reached when iteration t - 1 (and its check) ended
  const int row_ctas = dv.rows / kBlockRows;
  if (ev) TL_CUDA(cudaEventRecord(ev[0], s));
  if (rp.dot_form) { image_kernel<H><<<dv.slots / kBlockRows, kBlockRows, 0, s>>>(dv, cur, e_prev); rp.launches += 1; }
  if (rp.tc_form) launch_image_tc<H>(rp, s, cur, e_prev);
  if (overlap) {
    // The walk first: its CTAs announce themselves at once (griddepcontrol.launch_dependents), which lets the
    // repulsion grid - launched as a programmatic dependent, it has no data dependence on the walk - fill the
    // rest of the chip while the walk's few warps are already resident.  (Launched from two streams the
    // repulsion grid usually won the race for the SMs and the walk ran after it.)
    launch_spring<H>(rp, s, cur, e_prev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)rp.rep_ctas); cfg.blockDim = dim3((unsigned)rp.rep_threads);
    cfg.dynamicSmemBytes = rp.rep_smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (!rp.tc_form || dv.adaptive) {
      cfg.gridDim = dim3((unsigned)rp.f32_ctas); cfg.blockDim = dim3((unsigned)rp.f32_threads); cfg.dynamicSmemBytes = rp.f32_smem;
      TL_CUDA(cudaLaunchKernelEx(&cfg, (RepulseFn)rp.rep_fn, dv, cur, e_prev));
      rp.launches += dv.adaptive ? 1 : 0;
    }
    if (rp.tc_form) launch_repulse_tc<H>(rp, s, cur, true);
  } else {
    if (!rp.tc_form || dv.adaptive) {
      ((RepulseFn)rp.rep_fn)<<<rp.f32_ctas, rp.f32_threads, rp.f32_smem, s>>>(dv, cur, e_prev);
      rp.launches += dv.adaptive ? 1 : 0;
    }
    if (rp.tc_form) launch_repulse_tc<H>(rp, s, cur, false);
    if (ev) TL_CUDA(cudaEventRecord(ev[1], s));
    launch_spring<H>(rp, s, cur, e_prev);
    if (ev) TL_CUDA(cudaEventRecord(ev[2], s));
  }
  const unsigned e_iter = ++rp.epoch;
  combine_kernel<H><<<row_ctas, kBlockRows, 0, s>>>(dv, rp.prm, cur, e_iter);
  if (ev) TL_CUDA(cudaEventRecord(ev[3], s));
  rp.launches += 3;
  if (check || fin) {
    const unsigned e_mae = ++rp.epoch;
    launch_mae<H>(rp, s, nxt, e_iter, e_mae);
    if (ev) TL_CUDA(cudaEventRecord(ev[4], s));
    ctl_kernel<<<1, 256, 0, s>>>(dv, rp.prm, check ? 1 : 0, fin ? 1 : 0, e_mae);
    if (ev) TL_CUDA(cudaEventRecord(ev[5], s));
    snap_kernel<<<148, 256, 0, s>>>(dv, nxt);
    if (ev) TL_CUDA(cudaEventRecord(ev[6], s));
    rp.launches += 3;
  }
  TL_CUDA(cudaGetLastError());
  rp.iter_launched = t + 1;
}

template <int H>
void launch_spring(const RowPlan& rp, cudaStream_t s, int cur, unsigned e_spring) {
  const int ctas = rp.dv.rows / kBlockRows;
  if (rp.deep_ring) spring_kernel<H, 8><<<ctas, kBlockRows, kWalkSmem<H, 8>, s>>>(rp.dv, rp.prm, cur, e_spring);
  else spring_kernel<H, 4><<<ctas, kBlockRows, kWalkSmem<H, 4>, s>>>(rp.dv, rp.prm, cur, e_spring);
}
template <int H>
void launch_mae(const RowPlan& rp, cudaStream_t s, int nxt, unsigned e_spring, unsigned e_mae) {
  const int ctas = rp.dv.rows / kBlockRows;
  if (rp.deep_ring) mae_kernel<H, 8><<<ctas, kBlockRows, kWalkSmem<H, 8>, s>>>(rp.dv, nxt, e_spring, e_mae);
  else mae_kernel<H, 4><<<ctas, kBlockRows, kWalkSmem<H, 4>, s>>>(rp.dv, nxt, e_spring, e_mae);
}

void launch_one(RowPlan& rp, cudaStream_t s, cudaEvent_t* ev = nullptr) {
  dispatch_h((rp.dv.D + 1) / 2, [&](auto h) { launch_iteration<decltype(h)::value>(rp, s, ev, rp.overlap && !ev); });
}

// Shapes of the repulsion kernel: {rows per thread, partner unroll, register cap, cube on the SFU}.
// 0 is the production shape; the others stay selectable (TOPOLOW_REP_VARIANT) for measurements.
// Variant 0 = policy: 11 (rowblock_tc2.cuh: distances and accumulation on the tensor cores) from ndim 5 on, 5 (the difference
// form) below; 10 (rowblock_tc.cuh: distances only on the tensor cores) stays selectable.
// The FP32 inner-product forms (4, 6-9) are kept selectable as the measured record of why they are not used (B200, cfg4 /
// cfg3 repulsion pass in ms; difference form 16.8 / 0.23, tensor-core form 13.4 / 0.19):
//   6  one partner per trip, no spill          16.6 / 0.30   issue port 79 % busy (a packed instruction holds it 2 cycles):
//                                                            two dependency chains per warp cannot hide the MUFU chain
//   4  two partners per trip, 128 registers    19.1 / 0.31   needs 159 registers; capped, ptxas spills W / sp / thr and shuffles p
//   7  two partners per trip, 168 registers    18.1 / 0.30   3 warps per scheduler
//   8  as 6, power by series (1 MUFU)          17.5 / 0.27   +6 FMA-pipe slots cost more than 2 MUFU saved
//   9  as 4, power by series                   20.9 / 0.26
// The difference form issues 24 packed + 8.75 other instructions per interaction (56.75 port cycles) and runs at 60.6
// cycles: it is bound by the issue port, not by the FMA pipe's 51 cycles, and the inner-product form's 49 port cycles only
// pay with four chains in flight, which do not fit beside 64 registers of own rows and accumulators.
template <int H>
void repulse_variant(int v, RepulseFn* fn, int* threads, bool* dot, int* stage) {
  *dot = false;
  *stage = kStageJ;
  *threads = kRowTile / 2;
  switch (v) {
    case 1: *fn = repulse_kernel<H, 2, 2, 128, false>; break;
    case 2: *fn = repulse_kernel<H, 2, 2, 120, false>; break;
    case 3: *fn = repulse_kernel<H, 2, 2, 120, true>; break;
    case 4: *fn = repulse_dot_kernel<H, 2, 2, 128>; *dot = true; break;
    case 6: *fn = repulse_dot_kernel<H, 2, 1, 128>; *dot = true; break;
    case 7: *fn = repulse_dot_kernel<H, 2, 2, 168>; *dot = true; break;
    case 8: *fn = repulse_dot_kernel<H, 2, 1, 128, kStageJ, true>; *dot = true; break;
    case 9: *fn = repulse_dot_kernel<H, 2, 2, 128, kStageJ, true>; *dot = true; break;
    default: *fn = repulse_kernel<H, 2, 2, 128, true>; break;
  }
}
template <int H>
void configure_repulse(RowPlan& rp, int sms) {
  const char* ev = std::getenv("TOPOLOW_REP_VARIANT");
  RepulseFn fn; int threads;
  int stage = kStageJ;
  int variant = ev ? std::atoi(ev) : 0;
  // policy, measured on B200 (ms per repulsion pass, forms 11 / 10 / 5): ndim 16, 100k: 7.5 / 13.4 / 16.8; ndim 12, 40k: 1.24 / 1.82 /
  // 2.32; ndim 10, 10k: 0.16 / 0.19 / 0.23; ndim 6, 30k: 0.69 / 0.84 / 0.75.  The two-GEMM form costs the same at every ndim (K is
  // padded to 16 coordinates); the difference form's cost falls with ndim and takes over below ndim 5.
  if (variant == 0) variant = H >= 3 ? 11 : 5;
  rp.rep_form = variant;
  if (variant == 10 || variant == 11) {          // distances (10) / distances and accumulation (11) on the tensor cores
    rp.tc_form = true;
    rp.tc2_form = variant == 11;
    rp.dv.rparts = (rp.tc2_form ? 3 : 2) * rp.dv.chunks;
    if (rp.tc2_form) {
      TL_CUDA(cudaFuncSetAttribute(repulse_tc2_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem::kTotal));
      TL_CUDA(cudaFuncSetAttribute(repulse_tc2_kernel<H>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    TL_CUDA(cudaFuncSetAttribute(repulse_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::kTotal));
    TL_CUDA(cudaFuncSetAttribute(repulse_tc_kernel<H>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    TL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, repulse_tc_kernel<H>, kTcThreads, TcSmem::kTotal));
    if (std::getenv("TOPOLOW_DEBUG")) {
      int sm_smem = 0, blk_smem = 0, regs_sm = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
      cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, rp.device);
      cudaDeviceGetAttribute(&blk_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, rp.device);
      cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, rp.device);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, repulse_tc_kernel<H>, kTcThreads, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, repulse_tc_kernel<H>, kTcThreads, 65536);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, repulse_tc_kernel<H>, 256, TcSmem::kTotal);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, repulse_tc_kernel<H>, kTcThreads, 90000);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, repulse_tc_kernel<H>);
      std::fprintf(stderr, "[topolow] repulse_tc: %d CTAs per SM by the occupancy API at %d bytes (0 B: %d, 64 KB: %d, 90000 B: %d, 256 threads: %d); "
                   "SM smem %d, block optin %d, regs/SM %d; kernel regs %d static smem %d maxdyn %d carveout %d\n",
                   per_sm, (int)TcSmem::kTotal, o0, o1, o3, o2, sm_smem, blk_smem, regs_sm, fa.numRegs, (int)fa.sharedSizeBytes,
                   fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout);
    }
    // The occupancy API answers 1 for a kernel that allocates tensor memory (it cannot know how many of the 512 columns
    // tcgen05.alloc will ask for); registers, shared memory and the 256 columns a CTA takes all allow two, and the work
    // split does not depend on how many CTAs are resident at once.
    per_sm = 2;
    if (const char* ec = std::getenv("TOPOLOW_TC_CTAS")) per_sm = std::max(1, std::atoi(ec));
    const long long pairs = (long long)(rp.dv.rows / kRowTile) * rp.dv.chunks;
    rp.tc_cpi = (int)std::max<long long>(1, std::min<long long>(rp.dv.chunks, pairs / (6ll * sms * per_sm)));
    if (const char* ec = std::getenv("TOPOLOW_TC_CPI")) rp.tc_cpi = std::max(1, std::atoi(ec));
    const long long items = (long long)(rp.dv.rows / kRowTile) * ((rp.dv.chunks + rp.tc_cpi - 1) / rp.tc_cpi);
    rp.rep_ctas = (int)std::min<long long>((long long)sms * per_sm, std::max<long long>(items, 1));
    rp.rep_threads = kTcThreads; rp.rep_smem = TcSmem::kTotal; rp.rep_fn = nullptr;
  }
  repulse_variant<H>(rp.tc_form ? 5 : variant, &fn, &threads, &rp.dot_form, &stage);
  const size_t smem = rp.dot_form ? (size_t)kStages * stage * Img<H>::kStride * sizeof(float) + (size_t)2 * H * threads * sizeof(float2)
                                  : (size_t)kStages * kStageJ * Row<H>::kStride * sizeof(float);
  if (smem > 48 * 1024) TL_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
  if (per_sm < 1) per_sm = 1;
  if (const char* ec = std::getenv("TOPOLOW_REP_CTAS")) per_sm = std::max(1, std::atoi(ec));
  const long long items = (long long)(rp.dv.rows / kRowTile) * rp.dv.chunks;
  rp.f32_ctas = (int)std::min<long long>((long long)sms * per_sm, std::max<long long>(items, 1));
  rp.f32_smem = smem; rp.f32_threads = threads;
  rp.rep_fn = (void*)fn;
  if (!rp.tc_form) { rp.rep_ctas = rp.f32_ctas; rp.rep_smem = smem; rp.rep_threads = threads; }
  // The policy's tensor form is adaptive: the device picks the form of every iteration (image_tc_kernel's probe) and
  // the host launches both kernels.  A form asked for by name (TOPOLOW_REP_VARIANT) runs unconditionally.
  rp.dv.adaptive = (rp.tc_form && !ev) ? 1 : 0;
  if (const char* ea = std::getenv("TOPOLOW_ADAPTIVE")) rp.dv.adaptive = (rp.tc_form && std::atoi(ea) != 0) ? 1 : 0;
  rp.dv.series = rp.dv.adaptive;     // not adaptive: the series form only on request (it is exact only for pairs farther than 0.1 apart)
  if (const char* es = std::getenv("TOPOLOW_TC_SERIES")) rp.dv.series = std::atoi(es) != 0 ? 1 : 0;
  rp.dv.probe_k = (int)std::min<long long>(64, std::max<long long>(4, (262144 + rp.dv.n - 1) / std::max(rp.dv.n, 1)));
  // Break-even, measured at cfg4 (ndim 16): the fix-ups cost about 4.7 s x (near fraction) per pass over 1e10 one-sided pairs,
  // the difference form 9 ms more than the tensor form: near fraction 2e-3; half of that is the limit, scaled with ndim
  // (the difference form's cost falls with ndim, the tensor form's does not).
  const double p_lim = 1.25e-4 * H;
  if (const char* el = std::getenv("TOPOLOW_NEAR_LIMIT")) { rp.near_limit = std::atof(el); } else rp.near_limit = p_lim;
  rp.dv.probe_limit = (unsigned)((double)rp.dv.probe_k * (double)rp.dv.n * rp.near_limit);
  rp.sm_count = sms;
  rp.overlap = true;   // measured on B200 at cfg4: one rank of 8 2.61 -> 2.38 ms, of 2 9.11 -> 8.99, one GPU 18.29 -> 18.11; cfg3 0.47 -> 0.29
  if (const char* eo = std::getenv("TOPOLOW_OVERLAP")) rp.overlap = std::atoi(eo) != 0;
  // spring / MAE walk: the deep ring needs more than the default 48 KB of dynamic shared memory
  rp.deep_ring = rp.dv.rows / kBlockRows <= 2 * sms;
  if (const char* er = std::getenv("TOPOLOW_DEEP_RING")) rp.deep_ring = std::atoi(er) != 0;
  TL_CUDA(cudaFuncSetAttribute(spring_kernel<H, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWalkSmem<H, 8>));
  TL_CUDA(cudaFuncSetAttribute(mae_kernel<H, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWalkSmem<H, 8>));
}

}  // namespace

RowPlan* row_create(const topolow_problem& pb, const topolow_params& pr, int rank, int n_ranks) {
  validate(pb, pr);
  if (n_ranks < 1 || n_ranks > kMaxShards || rank < 0 || rank >= n_ranks) throw BadArg("rank / n_ranks out of range");
  if (pb.ndim > 16) throw BadArg("row-block mode supports ndim <= 16");
  if (pb.n >= (1ll << 30)) throw BadArg("n out of range");
  TL_CUDA(cudaSetDevice(pr.device));
  keep_pool_memory(pr.device);
  int sm_count = 0;
  TL_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, pr.device));   // (cudaGetDeviceProperties costs milliseconds)
  std::unique_ptr<RowPlan> rp(new RowPlan());
  rp->device = pr.device;
  rp->n = pb.n; rp->E = pb.n_edges;
  rp->prm = FitParams{pr.n_iter, pr.k0, pr.cooling_rate, pr.c_repulsion, pr.relative_epsilon, pr.convergence_window,
                      pr.convergence_check_freq};
  RowDev& dv = rp->dv;
  dv.n = (int)pb.n; dv.D = pb.ndim; dv.Dp = (pb.ndim + 3) / 4 * 4;
  dv.G = n_ranks; dv.rank = rank;
  const int tiles = (int)((pb.n + kRowTile - 1) / kRowTile);
  dv.slots = tiles * kRowTile;
  const int t0 = (int)((long long)tiles * rank / n_ranks), t1 = (int)((long long)tiles * (rank + 1) / n_ranks);
  dv.row0 = t0 * kRowTile; dv.rows = (t1 - t0) * kRowTile;
  if (tiles < n_ranks) throw BadArg("too few points for this many ranks (every rank needs at least 256 rows)");
  dv.chunks = (int)((pb.n + kChunk - 1) / kChunk);
  dv.rparts = dv.chunks;
  dv.cap_rows = align_up((size_t)dv.slots, kChunk);
  dv.pairs_per_iter = (unsigned long long)pb.n * (unsigned long long)(pb.n - 1) / 2ull;
  dv.seed = pr.seed;
  dv.no_wait = std::getenv("TOPOLOW_IGNORE_PEERS") ? 1 : 0;
  TL_CUDA(cudaStreamCreate(&rp->stream));
  TL_CUDA(cudaEventCreate(&rp->ev0));
  TL_CUDA(cudaEventCreate(&rp->ev1));
  PhaseTimer pt(rp->stream);
  // relabelling: a fixed pseudo-random permutation (decorrelates the caller's row order from the row
  // blocks and from the order a row visits its partners in)
  rp->slot_of_point = random_permutation(pb.n, kRowLayoutSeed);
  pt.mark("rows: relabelling");

  // ---- shared block ----
  const size_t pos_bytes = 2 * dv.cap_rows * dv.Dp * sizeof(float);
  rp->off_red = align_up(pos_bytes, 256);
  rp->off_flags = align_up(rp->off_red + (size_t)(dv.slots / kBlockRows) * 4 * sizeof(double), 256);
  rp->shared_bytes = align_up(rp->off_flags + (size_t)kMaxShards * kFlagStride * sizeof(unsigned), 256);
  TL_CUDA(cudaMalloc((void**)&rp->shared, rp->shared_bytes));
  TL_CUDA(cudaMemset(rp->shared, 0, rp->shared_bytes));
  set_self_pointers(*rp);
  pt.mark("rows: shared block");

  // ---- points ----
  {
    std::vector<float> hp(dv.cap_rows * dv.Dp, 0.f), hd(dv.slots, 0.f);
    for (int64_t i = 0; i < pb.n; ++i) {
      const size_t sl = rp->slot_of_point[i];
      for (int d = 0; d < pb.ndim; ++d) hp[sl * dv.Dp + d] = (float)pb.initial_positions[(size_t)d * pb.n + i];
      hd[sl] = (float)((double)pb.degrees[i] + 1.0);
    }
    pool_alloc(rp->best, hp.size() * sizeof(float));
    pool_alloc(rp->dp1, hd.size() * sizeof(float));
    pool_alloc(rp->rpart, (size_t)3 * dv.chunks * std::max(dv.rows, 1) * dv.Dp * sizeof(float));
    pool_alloc(rp->tc_buf, dv.cap_rows * (size_t)(16 + 16 + 4 + 4 + 16 + 32 + 16) * sizeof(float));
    pool_alloc(rp->xs, (size_t)std::max(dv.rows, 1) * dv.Dp * sizeof(float));
    pool_alloc(rp->img, dv.cap_rows * (size_t)(((dv.D + 1) / 2 * 2 + 2 + 3) / 4 * 4) * sizeof(float));   // Img<H>::kStride floats per row
    pool_alloc(rp->hmax, 2 * sizeof(float));
    pool_alloc(rp->state, sizeof(FitState));
    pool_alloc(rp->trace, sizeof(double) * std::max(pr.n_iter, 1));
    pool_alloc(rp->counters, 8 * sizeof(unsigned));
    pool_ready();
    TL_CUDA(cudaMemcpy(rp->shared, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy(rp->best, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy(rp->dp1, hd.data(), hd.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<double> tr(std::max(pr.n_iter, 1), NAN);
    TL_CUDA(cudaMemcpy(rp->trace, tr.data(), tr.size() * sizeof(double), cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemset(rp->counters, 0, 8 * sizeof(unsigned)));
    TL_CUDA(cudaMemset(rp->hmax, 0, 2 * sizeof(float)));
    // centre of the inner-product form: the centroid of the initial positions, summed in point order (the same on
    // every rank, fixed for the fit; the map stays about where it starts, and a map that wanders only makes more
    // pairs take the difference form)
    for (int d = 0; d < 16; ++d) dv.centre[d] = 0.f;
    for (int d = 0; d < pb.ndim; ++d) {
      double sum = 0.0;
      for (int64_t i = 0; i < pb.n; ++i) sum += pb.initial_positions[(size_t)d * pb.n + i];
      const double c = sum / (double)pb.n;
      dv.centre[d] = std::isfinite(c) ? (float)c : 0.f;
    }
    FitState st;
    state_init(st, rp->prm);
    TL_CUDA(cudaMemcpy(rp->state, &st, sizeof st, cudaMemcpyHostToDevice));
  }
  pt.mark("rows: points");
  dv.best = rp->best; dv.dp1 = rp->dp1; dv.rpart = rp->rpart; dv.xs = rp->xs; dv.img = rp->img; dv.hmax = rp->hmax;
  dv.state = rp->state; dv.trace = rp->trace;
  dv.counters = rp->counters;
  TL_CUDA(cudaHostAlloc((void**)&rp->h_flag, 2 * sizeof(int), cudaHostAllocMapped));
  rp->h_flag[0] = 0; rp->h_flag[1] = 0;
  TL_CUDA(cudaHostGetDevicePointer((void**)&rp->d_flag, (void*)rp->h_flag, 0));
  dv.host_flag = rp->d_flag;

  if (dv.rows > 0) build_records(*rp, pb);
  if (pb.n_holdout > 0) {
    if (!pb.holdout_i || !pb.holdout_j || !pb.holdout_truth) throw BadArg("hold-out arrays are required when n_holdout > 0");
    rp->n_holdout = pb.n_holdout;
    rp->hold_i.assign(pb.holdout_i, pb.holdout_i + pb.n_holdout);
    rp->hold_j.assign(pb.holdout_j, pb.holdout_j + pb.n_holdout);
    rp->hold_truth.assign(pb.holdout_truth, pb.holdout_truth + pb.n_holdout);
  }
  pt.mark("rows: records + hold-out");
  if (dv.rows > 0) dispatch_h((dv.D + 1) / 2, [&](auto h) { configure_repulse<decltype(h)::value>(*rp, sm_count); });
  pt.mark("rows: kernel attributes");
  TL_CUDA(cudaDeviceSynchronize());
  pt.mark("rows: device synchronize");
  rp->attached = (n_ranks == 1);
  return rp.release();
}

void row_destroy(RowPlan* rp) { delete rp; }

size_t row_handle_bytes() { return sizeof(RowHandle); }

void row_export(RowPlan& rp, void* blob) {
  DeviceScope on(rp.device);
  RowHandle h{};
  TL_CUDA(cudaIpcGetMemHandle(&h.mem, rp.shared));
  h.rank = rp.dv.rank; h.n_ranks = rp.dv.G; h.slots = rp.dv.slots; h.Dp = rp.dv.Dp; h.layout = rp.shared_bytes;
  std::memcpy(blob, &h, sizeof h);
}

void row_attach(RowPlan& rp, const void* blobs, int n_blobs) {
  DeviceScope on(rp.device);
  if (n_blobs != rp.dv.G) throw BadArg("one handle per rank is required");
  const RowHandle* hs = static_cast<const RowHandle*>(blobs);
  for (int q = 0; q < n_blobs; ++q) {
    RowHandle h;
    std::memcpy(&h, hs + q, sizeof h);
    if (h.rank != q || h.n_ranks != rp.dv.G || h.slots != rp.dv.slots || h.Dp != rp.dv.Dp || h.layout != rp.shared_bytes)
      throw BadArg("the ranks of a sharded map must be created from the same problem");
    if (q == rp.dv.rank) continue;
    void* base = nullptr;
    TL_CUDA(cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess));
    rp.peer_base[q] = base; rp.peer_ipc[q] = true;
    set_peer_pointers(rp, q, static_cast<char*>(base));
  }
  rp.attached = true;
}

void row_attach_local(RowPlan* const* plans, int n) {
  if (n < 1 || n > kMaxShards) throw BadArg("number of replicas out of range");
  for (int a = 0; a < n; ++a) {
    RowPlan& pa = *plans[a];
    if (pa.dv.G != n || pa.dv.rank != a) throw BadArg("replicas must be passed in rank order");
    if (pa.dv.slots != plans[0]->dv.slots || pa.dv.Dp != plans[0]->dv.Dp) throw BadArg("the ranks of a sharded map must be created from the same problem");
    DeviceScope on(pa.device);
    for (int b = 0; b < n; ++b) {
      if (a == b) continue;
      RowPlan& pb = *plans[b];
      if (pb.device != pa.device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(pb.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TL_CUDA(e);
        cudaGetLastError();
      }
      pa.peer_base[b] = pb.shared; pa.peer_ipc[b] = false;
      set_peer_pointers(pa, b, pb.shared);
    }
    pa.attached = true;
  }
}

static void check_runnable(const RowPlan& rp) {
  if (!rp.attached) throw BadArg("the replicas of a sharded map must be attached before it runs");
}

double row_run(RowPlan& rp, int n_iters, cudaStream_t stream_in, topolow_interrupt_fn poll, void* user, bool* interrupted) {
  check_runnable(rp);
  TL_CUDA(cudaSetDevice(rp.device));
  cudaStream_t s = stream_in ? stream_in : rp.stream;
  TL_CUDA(cudaEventRecord(rp.ev0, s));
  int left = std::min(n_iters, std::max(0, rp.prm.n_iter - rp.iter_launched));
  int since_sync = 0;
  while (left > 0) {
    if (rp.h_flag[0]) break;
    if (poll && since_sync == 0 && poll(user)) { if (interrupted) *interrupted = true; break; }
    launch_one(rp, s);
    --left;
    if (++since_sync >= 50) {             // the reference's interrupt interval (src/optimization.cpp:364)
      since_sync = 0;
      if (poll) TL_CUDA(cudaStreamSynchronize(s));
    }
  }
  TL_CUDA(cudaEventRecord(rp.ev1, s));
  TL_CUDA(cudaEventSynchronize(rp.ev1));
  float ms = 0.f;
  TL_CUDA(cudaEventElapsedTime(&ms, rp.ev0, rp.ev1));
  rp.total_ms += ms;
  return ms;
}

double row_run_local(RowPlan* const* plans, int n, int n_iters) {
  for (int a = 0; a < n; ++a) check_runnable(*plans[a]);
  bool same_device = true;
  for (int a = 1; a < n; ++a) same_device = same_device && plans[a]->device == plans[0]->device;
  RowPlan& p0 = *plans[0];
  int left = std::min(n_iters, std::max(0, p0.prm.n_iter - p0.iter_launched));
  TL_CUDA(cudaSetDevice(p0.device));
  TL_CUDA(cudaEventRecord(p0.ev0, p0.stream));
  if (same_device) {
    // lock step on ONE stream: rank after rank, phase after phase, so that every wait finds its epoch reached
    cudaStream_t s = p0.stream;
    for (int a = 0; a < n; ++a) TL_CUDA(cudaStreamSynchronize(plans[a]->stream));
    while (left > 0) {
      if (p0.h_flag[0]) break;
      const int t = p0.iter_launched;
      const int cur = t & 1, nxt = cur ^ 1;
      const bool check = is_check_iter(t, p0.prm), fin = ((t + 1) % 10 == 0);
      for (int a = 0; a < n; ++a) {
        RowPlan& rp = *plans[a];
        dispatch_h((rp.dv.D + 1) / 2, [&](auto h) {
          constexpr int H = decltype(h)::value;
          const unsigned e_prev = rp.epoch;
          if (rp.dot_form) { image_kernel<H><<<rp.dv.slots / kBlockRows, kBlockRows, 0, s>>>(rp.dv, cur, e_prev); rp.launches += 1; }
          if (rp.tc_form) launch_image_tc<H>(rp, s, cur, e_prev);
          if (!rp.tc_form || rp.dv.adaptive) {
            ((RepulseFn)rp.rep_fn)<<<rp.f32_ctas, rp.f32_threads, rp.f32_smem, s>>>(rp.dv, cur, e_prev);
            rp.launches += rp.dv.adaptive ? 1 : 0;
          }
          if (rp.tc_form) launch_repulse_tc<H>(rp, s, cur, false);
          launch_spring<H>(rp, s, cur, e_prev);
          combine_kernel<H><<<rp.dv.rows / kBlockRows, kBlockRows, 0, s>>>(rp.dv, rp.prm, cur, ++rp.epoch);
        });
        rp.launches += 3; rp.iter_launched = t + 1;
      }
      if (check || fin) {
        for (int a = 0; a < n; ++a) {
          RowPlan& rp = *plans[a];
          dispatch_h((rp.dv.D + 1) / 2, [&](auto h) {
            constexpr int H = decltype(h)::value;
            const unsigned e_iter = rp.epoch;
            launch_mae<H>(rp, s, nxt, e_iter, ++rp.epoch);
          });
          rp.launches += 1;
        }
        for (int a = 0; a < n; ++a) {
          RowPlan& rp = *plans[a];
          ctl_kernel<<<1, 256, 0, s>>>(rp.dv, rp.prm, check ? 1 : 0, fin ? 1 : 0, rp.epoch);
          snap_kernel<<<148, 256, 0, s>>>(rp.dv, nxt);
          rp.launches += 2;
        }
      }
      TL_CUDA(cudaGetLastError());
      --left;
      if ((t + 1) % 50 == 0) TL_CUDA(cudaStreamSynchronize(s));
    }
    TL_CUDA(cudaEventRecord(p0.ev1, s));
  } else {
    while (left > 0) {
      if (p0.h_flag[0]) break;
      for (int a = 0; a < n; ++a) {
        TL_CUDA(cudaSetDevice(plans[a]->device));
        launch_one(*plans[a], plans[a]->stream);
      }
      --left;
    }
    for (int a = 1; a < n; ++a) { TL_CUDA(cudaSetDevice(plans[a]->device)); TL_CUDA(cudaStreamSynchronize(plans[a]->stream)); }
    TL_CUDA(cudaSetDevice(p0.device));
    TL_CUDA(cudaEventRecord(p0.ev1, p0.stream));
  }
  TL_CUDA(cudaEventSynchronize(p0.ev1));
  float ms = 0.f;
  TL_CUDA(cudaEventElapsedTime(&ms, p0.ev0, p0.ev1));
  for (int a = 0; a < n; ++a) plans[a]->total_ms += ms;
  return ms;
}

void row_time_kernels(RowPlan& rp, int n_iters, double* out, int cap) {
  check_runnable(rp);
  TL_CUDA(cudaSetDevice(rp.device));
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int done = 0;
  std::vector<EventGuard> evs(7);
  cudaEvent_t ev[7];
  for (int i = 0; i < 7; ++i) ev[i] = evs[i];
  int left = std::min(n_iters, std::max(0, rp.prm.n_iter - rp.iter_launched));
  while (left-- > 0 && !rp.h_flag[0]) {
    const int t = rp.iter_launched;
    const bool extra = is_check_iter(t, rp.prm) || ((t + 1) % 10 == 0);
    launch_one(rp, rp.stream, ev);     // in order on one stream: every kernel is timed alone
    TL_CUDA(cudaStreamSynchronize(rp.stream));
    float ms = 0.f;
    TL_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1])); acc[0] += ms;
    TL_CUDA(cudaEventElapsedTime(&ms, ev[1], ev[2])); acc[1] += ms;
    TL_CUDA(cudaEventElapsedTime(&ms, ev[2], ev[3])); acc[6] += ms;
    if (extra) {
      TL_CUDA(cudaEventElapsedTime(&ms, ev[3], ev[4])); acc[2] += ms;
      TL_CUDA(cudaEventElapsedTime(&ms, ev[4], ev[5])); acc[3] += ms;
      TL_CUDA(cudaEventElapsedTime(&ms, ev[5], ev[6])); acc[4] += ms;
      acc[5] += 1.0;
    }
    ++done;
  }
  const double v[7] = {done ? acc[0] / done : 0.0, done ? acc[1] / done : 0.0, acc[5] > 0 ? acc[2] / acc[5] : 0.0,
                       acc[5] > 0 ? acc[3] / acc[5] : 0.0, acc[5] > 0 ? acc[4] / acc[5] : 0.0, acc[5],
                       done ? acc[6] / done : 0.0};
  for (int i = 0; i < cap && i < 7; ++i) out[i] = v[i];
}

void row_result(RowPlan& rp, topolow_result& res, bool interrupted) {
  TL_CUDA(cudaSetDevice(rp.device));
  TL_CUDA(cudaStreamSynchronize(rp.stream));
  FitState st;
  TL_CUDA(cudaMemcpy(&st, rp.state, sizeof st, cudaMemcpyDeviceToHost));
  const RowDev& dv = rp.dv;
  if (res.positions) {
    std::vector<float> hp((size_t)dv.slots * dv.Dp);
    TL_CUDA(cudaMemcpy(hp.data(), rp.best, hp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < rp.n; ++i) {
      const size_t sl = rp.slot_of_point[i];
      for (int d = 0; d < dv.D; ++d) res.positions[(size_t)d * rp.n + i] = (double)hp[sl * dv.Dp + d];
    }
  }
  res.converged = st.converged;
  res.iterations = st.best_iter;     // src/optimization.cpp:368-381: always the best snapshot
  res.final_mae = st.best_mae;
  res.final_k = st.best_k;
  res.iterations_run = st.iter;
  res.pair_updates = (int64_t)st.pair_updates;
  res.device_ms = rp.total_ms;
  if (dv.adaptive) {
    unsigned c[8];
    TL_CUDA(cudaMemcpy(c, rp.counters, sizeof c, cudaMemcpyDeviceToHost));
    rp.tensor_iters = (int64_t)c[7];
    if (std::getenv("TOPOLOW_DEBUG"))
      std::fprintf(stderr, "[topolow] rows: %lld of %d iterations ran the tensor form (probe: %d partners per row, limit %u near pairs)\n",
                   (long long)rp.tensor_iters, st.iter, dv.probe_k, dv.probe_limit);
  }
  res.fail_iter = st.fail_iter;
  res.status = TOPOLOW_OK;
  res.message[0] = 0;
  res.holdout_sum_abs = 0.0; res.holdout_count = 0;
  if (st.status == 2) {
    res.status = TOPOLOW_ERR_NONFINITE;
    std::snprintf(res.message, sizeof res.message, "Numerical instability at iteration %d. Reduce k0 or c_repulsion.", st.fail_iter);
  } else if (st.status == 3) {
    res.status = TOPOLOW_ERR_CUDA;
    std::snprintf(res.message, sizeof res.message, "a peer replica of the sharded map did not arrive within 20 s");
  } else if (interrupted) {
    res.status = TOPOLOW_ERR_INTERRUPTED;
    std::snprintf(res.message, sizeof res.message, "interrupted");
  }
  if (res.status == TOPOLOW_OK && rp.n_holdout > 0 && res.positions &&
      topolow_holdout_errors(res.positions, rp.n, dv.D, rp.n_holdout, rp.hold_i.data(), rp.hold_j.data(), rp.hold_truth.data(),
                             &res.holdout_sum_abs, &res.holdout_count, rp.device) != TOPOLOW_OK)
    throw BadArg("hold-out cells out of range");
  if (res.trace_mae && rp.prm.n_iter > 0)
    TL_CUDA(cudaMemcpy(res.trace_mae, rp.trace, sizeof(double) * rp.prm.n_iter, cudaMemcpyDeviceToHost));
}

void row_info(const RowPlan& rp, int64_t* out, int cap) {
  const RowDev& dv = rp.dv;
  const int64_t v[18] = {dv.slots, dv.D, dv.Dp, dv.G, dv.rank, dv.row0, dv.rows, dv.chunks, rp.n_recs, rp.n_mrecs, rp.launches,
                         rp.h_flag ? rp.h_flag[1] : 0, rp.h_flag ? rp.h_flag[0] : 0,
                         (int64_t)(dv.G - 1) * dv.rows * dv.Dp * 4,   // position bytes this rank stores into its peers per iteration
                         (int64_t)(dv.rows / kRowTile) * dv.chunks, rp.rep_ctas, rp.rep_form, rp.tensor_iters};
  for (int i = 0; i < cap && i < 18; ++i) out[i] = v[i];
}

}  // namespace tl

// ---------------------------------------------------------------------------------------------------
// C ABI (include/topolow_b200.h, topolow_shard_*)
// ---------------------------------------------------------------------------------------------------
struct topolow_shard { tl::RowPlan* rp; };

using namespace tl;

namespace {
template <class F>
int guarded(char* message, int32_t message_len, F&& f) {
  try {
    f();
    return TOPOLOW_OK;
  } catch (const BadArg& e) {
    set_msg(message, message_len, e.what());
    return TOPOLOW_ERR_BAD_ARG;
  } catch (const std::invalid_argument& e) {
    set_msg(message, message_len, e.what());
    return TOPOLOW_ERR_BAD_ARG;
  } catch (const CudaError& e) {
    set_msg(message, message_len, e.what());
    cudaGetLastError();
    return TOPOLOW_ERR_CUDA;
  } catch (const std::exception& e) {
    set_msg(message, message_len, e.what());
    return TOPOLOW_ERR_BAD_ARG;
  }
}
}  // namespace

extern "C" {

int topolow_shard_create(const topolow_problem* problem, const topolow_params* params, int32_t rank, int32_t n_ranks,
                         topolow_shard** shard_out, char* message, int32_t message_len) {
  if (!problem || !params || !shard_out) return TOPOLOW_ERR_BAD_ARG;
  *shard_out = nullptr;
  if (problem->n < 2) { set_msg(message, message_len, "Need at least 2 points for embedding"); return TOPOLOW_ERR_TOO_FEW_POINTS; }
  return guarded(message, message_len, [&] {
    RowPlan* rp = row_create(*problem, *params, rank, n_ranks);
    *shard_out = new topolow_shard{rp};
  });
}
int64_t topolow_shard_handle_bytes(void) { return (int64_t)row_handle_bytes(); }
int topolow_shard_export(topolow_shard* shard, void* handle_out) {
  if (!shard || !handle_out) return TOPOLOW_ERR_BAD_ARG;
  return guarded(nullptr, 0, [&] { row_export(*shard->rp, handle_out); });
}
int topolow_shard_attach(topolow_shard* shard, const void* handles, int32_t n_handles, char* message, int32_t message_len) {
  if (!shard || !handles) return TOPOLOW_ERR_BAD_ARG;
  return guarded(message, message_len, [&] { row_attach(*shard->rp, handles, n_handles); });
}
int topolow_shard_attach_local(topolow_shard* const* shards, int32_t n, char* message, int32_t message_len) {
  if (!shards || n < 1 || n > kMaxShards) return TOPOLOW_ERR_BAD_ARG;
  return guarded(message, message_len, [&] {
    RowPlan* plans[kMaxShards];
    for (int a = 0; a < n; ++a) { if (!shards[a]) throw BadArg("null shard"); plans[a] = shards[a]->rp; }
    row_attach_local(plans, n);
  });
}
int topolow_shard_run(topolow_shard* shard, int32_t n_iters, void* stream, double* ms_out) {
  if (!shard) return TOPOLOW_ERR_BAD_ARG;
  return guarded(nullptr, 0, [&] {
    const double ms = row_run(*shard->rp, n_iters, (cudaStream_t)stream, nullptr, nullptr, nullptr);
    if (ms_out) *ms_out = ms;
  });
}
int topolow_shard_run_local(topolow_shard* const* shards, int32_t n, int32_t n_iters, double* ms_out) {
  if (!shards || n < 1 || n > kMaxShards) return TOPOLOW_ERR_BAD_ARG;
  return guarded(nullptr, 0, [&] {
    RowPlan* plans[kMaxShards];
    for (int a = 0; a < n; ++a) { if (!shards[a]) throw BadArg("null shard"); plans[a] = shards[a]->rp; }
    const double ms = row_run_local(plans, n, n_iters);
    if (ms_out) *ms_out = ms;
  });
}
int topolow_shard_time_kernels(topolow_shard* shard, int32_t n_iters, double* out, int32_t cap) {
  if (!shard || !out) return TOPOLOW_ERR_BAD_ARG;
  return guarded(nullptr, 0, [&] { row_time_kernels(*shard->rp, n_iters, out, cap); });
}
int topolow_shard_result(topolow_shard* shard, topolow_result* result) {
  if (!shard || !result) return TOPOLOW_ERR_BAD_ARG;
  const int rc = guarded(result->message, sizeof result->message, [&] { row_result(*shard->rp, *result, false); });
  if (rc != TOPOLOW_OK) { result->status = rc; return rc; }
  return result->status;
}
int topolow_shard_info(const topolow_shard* shard, int64_t* out, int32_t cap) {
  if (!shard || !out) return TOPOLOW_ERR_BAD_ARG;
  row_info(*shard->rp, out, cap);
  return TOPOLOW_OK;
}
// slot_of_point of the row-block layout (a pure function of n): lets a checker restate the order rows
// are grouped and partners are visited in.
int topolow_shard_slot_order(int64_t n, int32_t* slot_of_point_out) {
  if (n < 1 || !slot_of_point_out) return TOPOLOW_ERR_BAD_ARG;
  const std::vector<int32_t> p = random_permutation(n, kRowLayoutSeed);
  std::memcpy(slot_of_point_out, p.data(), (size_t)n * sizeof(int32_t));
  return TOPOLOW_OK;
}
void topolow_shard_destroy(topolow_shard* shard) {
  if (!shard) return;
  row_destroy(shard->rp);
  delete shard;
}

}  // extern "C"
