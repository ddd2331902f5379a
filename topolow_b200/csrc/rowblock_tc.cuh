// topolow_b200/csrc/rowblock_tc.cuh - included by rowblock.cu inside namespace tl::{anonymous}.
//
// Repulsion pass with the pair distances on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// The difference form of repulse_kernel is bound by the issue port: 24 packed FP32 instructions per interaction, 16 of
// them only to get |q - p|^2.  Here d^2 / 2 = |p|^2 / 2 + |q|^2 / 2 - p.q for a 128 x 64 block of (own row, partner) pairs is
// ONE small GEMM: D[128 x 64] = -A[128 x 24] B[64 x 24]^T with K = 16 coordinates (zero padded) + 4 columns that add the two
// half norms (A: -h_i, -h_i', -1, -1; B: 1, 1, h_j, h_j') + 4 zeros, in TF32 with the 3-pass split that keeps FP32
// accuracy (x = hi + lo, hi = x rounded to TF32: hi.hi + hi.lo + lo.hi; lo.lo is below 2^-22).  The CUDA cores are left
// with what is not a contraction: read d^2 / 2 from TMEM (tcgen05.ld, one register per pair), weight = (d + 0.01)^-3 on the
// special-function unit, and the accumulation A_i += w x_j, W_i += w (8 packed FMA) with partner rows broadcast from
// shared memory; the sum over partners is A_i - x_i W_i.  About 30 issue-port cycles per interaction instead of 57.
//
// Pairs whose inner-product distance cannot be trusted get weight 0 in the main loop and are re-done from the coordinate
// differences (the point itself, difference exactly 0, contributes nothing, as in the difference form).  The error of
// d^2 / 2 is about 7e-7 (h_i + h_j), h = |x|^2 / 2; "near" means d^2 / 2 < tau (h_i + h_j) with tau = 1e-3 (relative error of
// d^2 could pass 1e-3).  A near partner has |x_j| <= |x_i| + d, hence h_j <= 2 h_i + d^2, and the test
// d^2 / 2 < 3.01 tau h_i - a per-row constant - catches every such pair whatever the rest of the map looks like.
// In 16 dimensions that is the point itself and hardly anything else.
//
// Roles in a CTA of 9 warps: warp 8, one elected lane, is producer and MMA issuer (1-D bulk copies global -> shared
// signalled on mbarriers, tcgen05.mma, tcgen05.commit); warps 0-7 consume: warp w reads TMEM lanes 32 (w % 4) .. + 31
// (= rows of both 128-row tiles) and the column half w / 4 of every 64-partner stage.  Three rings: shared-memory stages
// (full / empty), two TMEM buffers (full / empty), and the A operand of the work item.
//
// Operand layout (K-major, no swizzle; in 16-byte units ((8, n), 2) : ((1, SBO), LBO)): "chunk planes" - for every group of
// four consecutive K values one plane [rows][4 floats]; SBO = 128 bytes (8 rows), LBO = rows x 16 bytes (next plane).
// image_tc_kernel writes the planes for ALL points once per iteration, so a stage or an A tile is a handful of
// contiguous copies.

constexpr int kTcSJ = 64;            // partners per stage = N of one MMA
constexpr int kTcStages = 4;
constexpr int kTcRows = 256;         // own rows per work item = two M = 128 tiles (= kRowTile)
constexpr int kTcThreads = 288;      // 8 consumer warps + 1 control warp
constexpr int kTcTmemCols = 256;     // 2 buffers x 2 tiles x 64 columns
constexpr unsigned kTcSpinLimit = 1u << 28;

struct TcSmem {                      // byte offsets inside the dynamic shared memory of a CTA
  static constexpr int kAHi = 0;                                   // 6 planes x 256 rows x 16 B
  static constexpr int kALo = kAHi + 6 * kTcRows * 16;             // 4 planes
  static constexpr int kStage0 = kALo + 4 * kTcRows * 16;
  static constexpr int kBHi = 0;                                   // inside a stage: 6 planes x 64 x 16 B
  static constexpr int kBLo = 6 * kTcSJ * 16;                      // 4 planes
  static constexpr int kRows32 = kBLo + 4 * kTcSJ * 16;            // 64 x 64 B
  static constexpr int kStageBytes = kRows32 + kTcSJ * 64;
  static constexpr int kBars = kStage0 + kTcStages * kStageBytes;  // mbarriers (8 B each)
  static constexpr int kNumBars = 1 + 2 * kTcStages + 4;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kStageTx = 4 * kTcSJ * 16 + kTcSJ * 16 + 4 * kTcSJ * 16 + kTcSJ * 64;   // bytes landing per stage
  static constexpr int kATx = 4 * kTcRows * 16 + kTcRows * 16 + 4 * kTcRows * 16;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
TL_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
TL_D void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
TL_D void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
TL_D void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Every calling thread polls.  A wait that does not end is a protocol bug: trap (the launch fails) instead of hanging the device.
TL_D void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > kTcSpinLimit) __trap();
  }
}
TL_D void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
TL_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TL_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
TL_D void tc_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
TL_D void tc_mma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
TL_D void tc_ld16(unsigned taddr, unsigned (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
TL_D void tc_ld8(unsigned taddr, unsigned (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
TL_D unsigned tc_ld1(unsigned taddr) {
  unsigned v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
TL_D void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 in bits [0,14), leading
// byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64)).
TL_D unsigned long long tc_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (unsigned long long)((smem_addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), A negated
// (1 << 13), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
constexpr unsigned kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((unsigned)(kTcSJ >> 3) << 17) | ((128u >> 4) << 24);

constexpr float kSeriesNearS = 0.005f;   // = kT2SeriesS (rowblock_tc2.cuh): the series form of the weights takes pairs closer than 0.1 as near pairs

TL_D float tf32_round(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// ---- the image: chunk planes of hi / lo, the norm columns, FP32 rows ---------------------------------
struct TcImage {
  float* xhi;     // [4][cap_rows][4]  coordinates - centre, rounded to TF32
  float* xlo;     // [4][cap_rows][4]  the remainder
  float* aug_a;   // [cap_rows][4]     (-h_hi, -h_lo, -1, -1)     h = |x|^2 / 2
  float* aug_b;   // [cap_rows][4]     (1, 1, h_hi, h_lo); padding rows: h = 1e30 (their weight underflows to 0)
  float* rows32;  // [cap_rows][16]    coordinates - centre, FP32
};

template <int H>
__global__ void __launch_bounds__(kBlockRows) image_tc_kernel(RowDev dv, TcImage im, int cur, unsigned epoch) {
  constexpr int Dp = Row<H>::kStride;
  if (__ldcg(&dv.state->stop)) return;
  if (!wait_epoch(dv, epoch)) { if (blockIdx.x == 0 && threadIdx.x == 0) peer_timeout(dv); return; }
  const size_t row = (size_t)blockIdx.x * kBlockRows + threadIdx.x;     // grid covers cap_rows
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = 0.f;
  const bool real = row < (size_t)dv.n;
  if (real) {
    float2 p[H];
    ld_point<H>(dv.pos[dv.rank] + ((size_t)cur * dv.cap_rows + row) * Dp, p);
#pragma unroll
    for (int k = 0; k < H; ++k) { x[2 * k] = p[k].x - dv.centre[2 * k]; x[2 * k + 1] = p[k].y - dv.centre[2 * k + 1]; }
  }
  float h = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) h = fmaf(x[k], x[k], h);
  h *= 0.5f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 hi, lo;
    hi.x = tf32_round(x[4 * c]); hi.y = tf32_round(x[4 * c + 1]); hi.z = tf32_round(x[4 * c + 2]); hi.w = tf32_round(x[4 * c + 3]);
    lo.x = x[4 * c] - hi.x; lo.y = x[4 * c + 1] - hi.y; lo.z = x[4 * c + 2] - hi.z; lo.w = x[4 * c + 3] - hi.w;
    reinterpret_cast<float4*>(im.xhi)[(size_t)c * dv.cap_rows + row] = hi;
    reinterpret_cast<float4*>(im.xlo)[(size_t)c * dv.cap_rows + row] = lo;
    reinterpret_cast<float4*>(im.rows32)[row * 4 + c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
  }
  const float hb = real ? h : 1e30f;
  const float h_hi = tf32_round(h), hb_hi = tf32_round(hb);
  reinterpret_cast<float4*>(im.aug_a)[row] = make_float4(-h_hi, -(h - h_hi), -1.f, -1.f);
  reinterpret_cast<float4*>(im.aug_b)[row] = make_float4(1.f, 1.f, hb_hi, hb - hb_hi);
  if (dv.adaptive) {
    // ---- which form runs this iteration ----
    // The tensor form re-does every pair whose distance is small against the point's distance from the centre (the
    // cancellation in |x|^2 / 2 + |y|^2 / 2 - x.y) in the difference form, one divergent call each: cheap in a map that has
    // spread out, ruinous in the first iterations of the reference's start (R/core.R:407-415: cumulative steps along the
    // diagonal, 3 % of all pairs are near: 136 ms instead of 7.5 at cfg4).  Every row tests a few pseudo-random partners
    // against the same criterion; the count is a function of the replica alone, so every rank of a sharded map decides
    // the same way and the result still does not depend on the rank count.
    unsigned near = 0;
    if (real) {
      const unsigned it = (unsigned)__ldcg(&dv.state->iter);
      for (int k = 0; k < dv.probe_k; ++k) {
        unsigned long long z = ((unsigned long long)row * 64ull + (unsigned long long)k) * 0x9E3779B97F4A7C15ull + (unsigned long long)it * 0xD1B54A32D192ED03ull + dv.seed;
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        const size_t j = (size_t)(z % (unsigned long long)dv.n);
        if (j == row) continue;
        float2 q[H];
        ld_point<H>(dv.pos[dv.rank] + ((size_t)cur * dv.cap_rows + j) * Dp, q);
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < H; ++c) {
          const float dx = q[c].x - dv.centre[2 * c] - x[2 * c], dy = q[c].y - dv.centre[2 * c + 1] - x[2 * c + 1];
          d2 = fmaf(dx, dx, d2); d2 = fmaf(dy, dy, d2);
        }
        near += (0.5f * d2 < 3.01e-3f * h_hi) ? 1u : 0u;
        near += (0.5f * d2 < fmaxf(3.01e-3f * h_hi, kSeriesNearS)) ? 0x10000u : 0u;   // near in the series form of the weights
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) near += __shfl_xor_sync(0xffffffffu, near, o);     // <= 64 x 32 per half: no carry
    if ((threadIdx.x & 31) == 0 && near) {
      if (near & 0xffffu) atomicAdd(&dv.counters[4], near & 0xffffu);
      atomicAdd(&dv.counters[3], near >> 16);
    }
    if (last_cta(&dv.counters[5]) && threadIdx.x == 0) {
      const unsigned found = atomicExch(&dv.counters[4], 0u), found_series = atomicExch(&dv.counters[3], 0u);
      // 2: tensor form, weights by the one-MUFU series; 1: tensor form, square root + reciprocal; 0: FP32 difference form
      const unsigned form = (dv.series && found_series <= dv.probe_limit) ? 2u : (found <= dv.probe_limit ? 1u : 0u);
      dv.counters[6] = form;
      dv.counters[7] += form ? 1u : 0u;
      __threadfence();
    }
  }
}

// A near pair, from the coordinate differences (the arithmetic of repulse_kernel).  Out of line and on accumulators in
// local memory: it must not cost the main loop registers, and it is rare in a map that has spread out (the point
// itself, once per row and iteration) but not in the first iterations of a tightly packed start.
template <int H>
__device__ __noinline__ void near_fix(const float* __restrict__ xi, const float* __restrict__ xj, float* __restrict__ facc) {
  float d2 = 0.f;
#pragma unroll
  for (int k = 0; k < 2 * H; ++k) { const float dl = xj[k] - __ldg(xi + k); d2 = fmaf(dl, dl, d2); }
  if (d2 > 0.f) {                                 // the point itself (and exact duplicates) contribute nothing
    const float ds = sqrt_approx(d2) + 0.01f;
    const float w = ex2_approx(-3.0f * lg2_approx(ds));
#pragma unroll
    for (int k = 0; k < 2 * H; ++k) facc[k] = fmaf(xj[k] - __ldg(xi + k), w, facc[k]);
  }
}

// ---- the pass ------------------------------------------------------------------------------------------
// Work items are (256-row tile, group of `cpi` partner chunks), dealt round-robin to the CTAs (equal cost, several per
// CTA).  The sums of a (row, chunk, column half) do not depend on who computed them or when: rpart has 2 entries per
// chunk, added by combine_kernel in a fixed order.
template <int H>
__global__ void __maxnreg__(80) repulse_tc_kernel(RowDev dv, TcImage im, int cur, int cpi) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  constexpr int Dp = Row<H>::kStride;
  if (__ldcg(&dv.state->stop)) return;
  if (dv.adaptive && !form_is_tensor(dv)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned sm0 = smem_u32(tc_smem);
  const unsigned bar_a_full = sm0 + TcSmem::kBars;
  auto bar_full = [&](int s) { return sm0 + TcSmem::kBars + 8u * (1 + s); };
  auto bar_empty = [&](int s) { return sm0 + TcSmem::kBars + 8u * (1 + kTcStages + s); };
  auto bar_tfull = [&](int b) { return sm0 + TcSmem::kBars + 8u * (1 + 2 * kTcStages + b); };
  auto bar_tempty = [&](int b) { return sm0 + TcSmem::kBars + 8u * (1 + 2 * kTcStages + 2 + b); };
  unsigned* tmem_ptr_s = reinterpret_cast<unsigned*>(tc_smem + TcSmem::kTmemPtr);

  // ---- set-up: zero the padding planes (K columns 20..23), barriers, TMEM ----
  for (int x = tid; x < kTcRows * 4; x += kTcThreads) reinterpret_cast<float*>(tc_smem + TcSmem::kAHi + 5 * kTcRows * 16)[x] = 0.f;
  for (int s = 0; s < kTcStages; ++s)
    for (int x = tid; x < kTcSJ * 4; x += kTcThreads)
      reinterpret_cast<float*>(tc_smem + TcSmem::kStage0 + s * TcSmem::kStageBytes + TcSmem::kBHi + 5 * kTcSJ * 16)[x] = 0.f;
  if (tid == 0) {
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 8); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the zeroed planes are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem0 = *tmem_ptr_s;

  const int tiles = dv.rows / kTcRows;
  const int groups = (dv.chunks + cpi - 1) / cpi;
  const long long items = (long long)tiles * groups;
  auto chunk_stages = [&](int c) { return (min(kChunk, dv.n - c * kChunk) + kTcSJ - 1) / kTcSJ; };

  if (warp == 8) {
    // =========================== producer + MMA issuer (one lane) ===========================
    // Two cursors walk the CTA's stages in the same order: the load cursor runs kTcStages - 1 stages ahead of the MMA
    // cursor, across chunk and item boundaries.  A load is issued AFTER the MMA of the current stage, so that waiting for
    // a free slot (= the consumers finishing the stage before) never delays an MMA: MMA(st + 1) runs while the consumers
    // work on stage st.
    if (lane == 0) {
      struct Cursor { long long item; int c, c_hi, s; size_t r0; bool valid; };
      auto open_item = [&](Cursor& cu) {
        cu.valid = cu.item < items;
        if (!cu.valid) return;
        const int tile = (int)(cu.item / groups), grp = (int)(cu.item % groups);
        cu.c = grp * cpi; cu.c_hi = min(dv.chunks, cu.c + cpi); cu.s = 0;
        cu.r0 = (size_t)dv.row0 + (size_t)tile * kTcRows;
      };
      auto advance = [&](Cursor& cu) {
        if (++cu.s < chunk_stages(cu.c)) return;
        cu.s = 0;
        if (++cu.c < cu.c_hi) return;
        cu.item += gridDim.x;
        open_item(cu);
      };
      Cursor ld, mm;
      ld.item = mm.item = blockIdx.x;
      open_item(ld); open_item(mm);
      unsigned load_g = 0, mma_g = 0, a_uses = 0;
      auto issue_load = [&]() {
        const int s = load_g % kTcStages;
        mbar_wait(bar_empty(s), ((load_g / kTcStages) & 1) ^ 1);
        const size_t j0 = (size_t)ld.c * kChunk + (size_t)ld.s * kTcSJ;
        const unsigned stg = sm0 + TcSmem::kStage0 + s * TcSmem::kStageBytes;
        mbar_expect_tx(bar_full(s), TcSmem::kStageTx);
        for (int c = 0; c < 4; ++c) {
          bulk_g2s(stg + TcSmem::kBHi + c * kTcSJ * 16, im.xhi + ((size_t)c * dv.cap_rows + j0) * 4, kTcSJ * 16, bar_full(s));
          bulk_g2s(stg + TcSmem::kBLo + c * kTcSJ * 16, im.xlo + ((size_t)c * dv.cap_rows + j0) * 4, kTcSJ * 16, bar_full(s));
        }
        bulk_g2s(stg + TcSmem::kBHi + 4 * kTcSJ * 16, im.aug_b + j0 * 4, kTcSJ * 16, bar_full(s));
        bulk_g2s(stg + TcSmem::kRows32, im.rows32 + j0 * 16, kTcSJ * 64, bar_full(s));
        ++load_g;
        advance(ld);
      };
      for (int x = 0; x < kTcStages - 1 && ld.valid; ++x) issue_load();
      while (mm.valid) {
        if (mm.s == 0 && mm.c % cpi == 0) {
          // a new item: the MMAs of the previous one read the A tile - they are done when its last stage was committed
          if (mma_g > 0) mbar_wait(bar_tfull((mma_g - 1) & 1), ((mma_g - 1) >> 1) & 1);
          mbar_expect_tx(bar_a_full, TcSmem::kATx);
          for (int c = 0; c < 4; ++c) {
            bulk_g2s(sm0 + TcSmem::kAHi + c * kTcRows * 16, im.xhi + ((size_t)c * dv.cap_rows + mm.r0) * 4, kTcRows * 16, bar_a_full);
            bulk_g2s(sm0 + TcSmem::kALo + c * kTcRows * 16, im.xlo + ((size_t)c * dv.cap_rows + mm.r0) * 4, kTcRows * 16, bar_a_full);
          }
          bulk_g2s(sm0 + TcSmem::kAHi + 4 * kTcRows * 16, im.aug_a + mm.r0 * 4, kTcRows * 16, bar_a_full);
          mbar_wait(bar_a_full, a_uses & 1);
          ++a_uses;
        }
        const int s = mma_g % kTcStages, b = mma_g & 1;
        mbar_wait(bar_full(s), (mma_g / kTcStages) & 1);
        mbar_wait(bar_tempty(b), ((mma_g >> 1) & 1) ^ 1);
        tc_fence_after();
        const unsigned stg = sm0 + TcSmem::kStage0 + s * TcSmem::kStageBytes;
#pragma unroll
        for (int m = 0; m < 2; ++m) {                       // the two 128-row tiles of the item
          const unsigned d = tmem0 + (unsigned)(b * 2 * kTcSJ + m * kTcSJ);
          const unsigned a_hi = sm0 + TcSmem::kAHi + m * 128 * 16, a_lo = sm0 + TcSmem::kALo + m * 128 * 16;
          const unsigned b_hi = stg + TcSmem::kBHi, b_lo = stg + TcSmem::kBLo;
          constexpr unsigned kLboA = kTcRows * 16, kLboB = kTcSJ * 16;
#pragma unroll
          for (int k = 0; k < 3; ++k)                       // hi . hi over all 24 K values (the norm columns live here)
            tc_mma_tf32(d, tc_desc(a_hi + k * 2 * kLboA, kLboA, 128), tc_desc(b_hi + k * 2 * kLboB, kLboB, 128), kTcIdesc, k > 0);
#pragma unroll
          for (int k = 0; k < 2; ++k) {                     // the cross terms over the 16 coordinates
            tc_mma_tf32(d, tc_desc(a_hi + k * 2 * kLboA, kLboA, 128), tc_desc(b_lo + k * 2 * kLboB, kLboB, 128), kTcIdesc, 1);
            tc_mma_tf32(d, tc_desc(a_lo + k * 2 * kLboA, kLboA, 128), tc_desc(b_hi + k * 2 * kLboB, kLboB, 128), kTcIdesc, 1);
          }
        }
        tc_commit(bar_tfull(b));
        ++mma_g;
        advance(mm);
        if (ld.valid) issue_load();
      }
    }
  } else {
    // ======================================= consumers =======================================
    const int quad = warp & 3, half = warp >> 2;
    const unsigned lane_sel = (unsigned)(quad * 32) << 16;
    unsigned g = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const int tile = (int)(item / groups), grp = (int)(item % groups);
      const int c_lo = grp * cpi, c_hi = min(dv.chunks, c_lo + cpi);
      const int lrow[2] = {tile * kTcRows + quad * 32 + lane, tile * kTcRows + 128 + quad * 32 + lane};
      float thr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) thr[r] = -3.01e-3f * __ldg(im.aug_a + (size_t)(dv.row0 + lrow[r]) * 4);   // aug_a.x = -h_i (TF32-rounded)
      for (int c = c_lo; c < c_hi; ++c) {
        float2 acc[2][H];
        float W[2] = {0.f, 0.f};
        float facc[32];                                         // sums of the near pairs (difference form), both rows
        bool fixed = false;
#pragma unroll
        for (int k = 0; k < 32; ++k) facc[k] = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < H; ++k) acc[r][k] = make_float2(0.f, 0.f);
        const int nst = chunk_stages(c);
        for (int st = 0; st < nst; ++st, ++g) {
          const int s = g % kTcStages, b = g & 1;
          mbar_wait(bar_full(s), (g / kTcStages) & 1);          // the FP32 rows of the stage are visible
          mbar_wait(bar_tfull(b), (g >> 1) & 1);                // its distances are in TMEM
          tc_fence_after();
          const float* __restrict__ q_s = reinterpret_cast<const float*>(tc_smem + TcSmem::kStage0 + s * TcSmem::kStageBytes + TcSmem::kRows32) + half * 32 * 16;
          const unsigned t_base = tmem0 + lane_sel + (unsigned)(b * 2 * kTcSJ + half * 32);
#pragma unroll 1
          for (int bt = 0; bt < 4; ++bt) {                     // 8 partners at a time
            unsigned s0[8], s1[8];
            tc_ld8(t_base + bt * 8, s0);
            tc_ld8(t_base + kTcSJ + bt * 8, s1);
            tc_wait_ld();
            bool flag = false;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              float2 q[H];
              ld_point<H>(q_s + (bt * 8 + jj) * 16, q);
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const float S = __uint_as_float(r == 0 ? s0[jj] : s1[jj]);          // d^2 / 2
                const bool nr = S < thr[r];
                // (d + 0.01)^-3 with d = sqrt(2 S): 2^(-1.5 - 3 log2(sqrt(S) + 0.01 / sqrt(2)))
                const float ds = sqrt_approx(fabsf(S)) + 0.00707106781f;       // |.|: a rounding-negative S of a near pair must not make a NaN
                float w = ex2_approx(fmaf(-3.0f, lg2_approx(ds), -1.5f));
                w = nr ? 0.f : w;
                flag |= nr;
                const float2 ww = make_float2(w, w);
#pragma unroll
                for (int k = 0; k < H; ++k) acc[r][k] = __ffma2_rn(q[k], ww, acc[r][k]);
                W[r] += w;
              }
            }
            if (__any_sync(0xffffffffu, flag)) {                // some pair of this batch is near: redo those from differences
              unsigned t0[8], t1[8];
              tc_ld8(t_base + bt * 8, t0);
              tc_ld8(t_base + kTcSJ + bt * 8, t1);
              tc_wait_ld();
              float sv[16];                                    // indexed in a loop: lives in local memory, on this path only
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) { sv[jj] = __uint_as_float(t0[jj]); sv[8 + jj] = __uint_as_float(t1[jj]); }
              if (flag) {
                fixed = true;
#pragma unroll 1
                for (int x = 0; x < 16; ++x) {
                  const int r = x >> 3, jj = x & 7;
                  if (sv[x] < (r ? thr[1] : thr[0]))
                    near_fix<H>(im.rows32 + (size_t)(dv.row0 + (r ? lrow[1] : lrow[0])) * 16, q_s + (bt * 8 + jj) * 16, facc + r * 16);
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(bar_tempty(b)); mbar_arrive(bar_empty(s)); }
        }
        // ---- the chunk's sums of this column half: sum_j w (x_j - x_i) = A - x_i W ----
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float2 xi[H], o[H];
          ld_point<H>(im.rows32 + (size_t)(dv.row0 + lrow[r]) * 16, xi);
#pragma unroll
          for (int k = 0; k < H; ++k) o[k] = make_float2(fmaf(-xi[k].x, W[r], acc[r][k].x), fmaf(-xi[k].y, W[r], acc[r][k].y));
          if (fixed) {
#pragma unroll 1
            for (int k = 0; k < H; ++k) { o[k].x += facc[r * 16 + 2 * k]; o[k].y += facc[r * 16 + 2 * k + 1]; }
          }
          st_point<H>(dv.rpart + ((size_t)(2 * c + half) * dv.rows + lrow[r]) * Dp, o);
        }
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(kTcTmemCols) : "memory");
}
