// topolow_b200/csrc/schedule.h
//
// The coloured-parallel schedule, as pure integer functions shared by the CUDA kernel
// (tilepass.cu) and the host enumerator (topolow_plan_enumerate).  It replaces
// std::shuffle(all_pairs) of src/optimization.cpp:196 by a structured permutation of the same
// N(N-1)/2 pairs: every unordered pair is visited exactly once per iteration, and pairs that
// run concurrently never share a point, so the whole iteration equals ONE sequential order of
// the reference's Gauss-Seidel loop (src/optimization.cpp:199-282) - the order
// topolow_plan_enumerate() writes out.
//
// Hierarchy (point -> tile of 32*P -> super-block of W tiles -> S = 2*G*m super-blocks):
//   level 2  round-robin tournament over super-blocks (circle method): S-1 cross rounds of
//            S/2 disjoint super-block pairs (one CTA task each) + 1 diagonal round; a barrier
//            across the G CTAs of the fit separates rounds.
//   level 1  inside a cross task (X,Y): W sub-rounds, warp w takes tile X[w] x tile Y[(w+v)%W];
//            inside a diagonal task: circle method over the W tiles of each super-block, then
//            every tile against itself; __syncthreads() separates sub-rounds.
//   level 0  tile x tile (32P x 32P): 32 steps of a systolic ring - lane a keeps A[a + 32p], p < P, in
//            registers, the B points travel through the lanes in groups (B[b + 32q]) by an odd stride
//            g, b = (a + s0 + g*i) mod 32; a step is P perfect matchings ("waves") one after the
//            other, wave w pairing A[a + 32p] with B[b + 32((p + w) mod P)].
//            tile x itself: the pairs among a lane's own P slots, then 31 XOR steps in which lane a
//            meets lane a^x, wave w pairing slot a + 32p with slot (a^x) + 32((w - p) mod P).
// Randomisation per iteration (stateless hashes of (seed, iter)): the tile -> super-block
// placement, the order of the rounds, the sub-round rotation and the ring's (s0, g).
#pragma once

#include "common.cuh"

namespace tl {


struct Geometry {
  int n;        // real points
  int T;        // tiles = ceil(n / (32 * P))
  int W;        // tiles per super-block == warps per CTA
  int G;        // CTAs cooperating on the fit
  int m;        // tasks per CTA per round
  int S;        // super-blocks = 2*G*m
  int D;        // dimensions
  int P;        // points of a tile held by one lane: tile = 32 * P points (1, 2 or 3)
  // The job this launch executes (a whole fit is kind 0 over all tiles with do_end = 1):
  int kind;     // 0: every pair among the tiles [t0, t0 + tc);  1: every pair between [t0, t0 + tc) and [y0, y0 + yc);
                // 2: no pair updates (end-of-iteration phase only)
  int t0, tc;   // first tile / number of tiles of the (X) range
  int y0, yc;   // the Y range of a kind-1 job
  int do_end;   // run the end-of-iteration phase (cooling, MAE, controller) after the pair updates
  int table;    // device only, set by the launcher: the CTA keeps this iteration's tile placement and round
                // order in a shared-memory table (perm_table_entries) instead of hashing them per task
  uint64_t seed;
};

// ---- stateless permutation of [0, size): 4-round Feistel + cycle walking ------------------
TL_HD uint32_t feistel_perm(uint32_t x, uint32_t size, uint64_t key) {
  if (size <= 1) return 0;
  int half = 1;
  while ((1u << (2 * half)) < size) ++half;
  const uint32_t mask = (1u << half) - 1u;
  do {
    uint32_t l = x >> half, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = (uint32_t)(mix64(((uint64_t)r << 8) + (uint64_t)round + key) & mask);
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (l << half) | r;
  } while (x >= size);
  return x;
}

TL_HD uint64_t iter_key(const Geometry& g, int iter, uint32_t salt) {
  return mix64(g.seed ^ mix64(((uint64_t)(uint32_t)iter << 32) | salt));
}

// Circle method: the q-th pair (q in [0, S/2)) of round rr in a tournament of S (even) teams.
TL_HD void circle_pair(int S, int rr, int q, int& x, int& y) {
  const int M = S - 1;
  if (q == 0) { x = M; y = rr; }
  else { x = (rr + q) % M; y = (rr - q + M) % M; }
}

// Tile held by tile-slot `slot` (super-block slot/W, position slot%W) in this iteration, or -1.
// kind 0: S*W slots over the tc tiles of the job.  kind 1: S/2 super-blocks per side, (S/2)*W slots
// over the tc (side 0) or yc (side 1) tiles of that side.
// (Every function below exists twice: `f_k(..., key)` takes the iteration key it needs - iter_key(g, iter,
// salt), which the kernel computes once per iteration - and `f(..., iter, ...)` derives it.)
constexpr int kIterKeys = 8;   // salts 0..7
TL_HD int tile_at_k(const Geometry& g, uint64_t key /* salt 1 (side 0) or 7 (side 1) */, int slot, int side = 0) {
  const uint32_t slots = (uint32_t)((g.kind == 0 ? g.S : g.S / 2) * g.W);
  const uint32_t t = feistel_perm((uint32_t)slot, slots, key ^ (uint64_t)(uint32_t)(side ? g.y0 : g.t0));
  const int count = side ? g.yc : g.tc, base = side ? g.y0 : g.t0;
  return (int)t < count ? base + (int)t : -1;
}
TL_HD int tile_at(const Geometry& g, int iter, int slot, int side = 0) {
  return tile_at_k(g, iter_key(g, iter, side == 0 ? 1u : 7u), slot, side);
}

// Number of cross rounds of the job and the r-th one executed in this iteration (a permutation).
TL_HD int cross_rounds(const Geometry& g) { return g.kind == 0 ? g.S - 1 : (g.kind == 1 ? g.S / 2 : 0); }
TL_HD int round_at_k(const Geometry& g, uint64_t key /* salt 2 */, int r) {
  return (int)feistel_perm((uint32_t)r, (uint32_t)cross_rounds(g), key ^ (uint64_t)(uint32_t)(g.t0 * 131 + g.y0));
}
TL_HD int round_at(const Geometry& g, int iter, int r) { return round_at_k(g, iter_key(g, iter, 2), r); }
// Entries of the per-iteration lookup table: tile_at of every side-0 slot, then of every side-1 slot
// (kind 1), then round_at of every round.
TL_HD int perm_table_side0(const Geometry& g) { return (g.kind == 0 ? g.S : g.S / 2) * g.W; }
TL_HD int perm_table_entries(const Geometry& g) { return g.S * g.W + cross_rounds(g); }
// The q-th CTA task of cross round rr: super-block X (side 0) against super-block Y (side 0 for a
// kind-0 job: circle method; side 1 for a kind-1 job: Latin square).
TL_HD void cross_task(const Geometry& g, int rr, int q, int& x, int& y) {
  if (g.kind == 0) circle_pair(g.S, rr, q, x, y);
  else { x = q; y = (q + rr) % (g.S / 2); }
}

struct RingParams { int s0, g, ginv; };
// Ring offset / stride for tile pair (ta, tb) (actual tile ids, order-insensitive).
TL_HD RingParams ring_params_k(uint64_t key /* salt 3 */, int ta, int tb) {
  const int lo = ta < tb ? ta : tb, hi = ta < tb ? tb : ta;
  const uint64_t h = mix64(key ^ (((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi));
  RingParams p;
  p.s0 = (int)(h & 31);
  p.g = (int)((h >> 5) & 15) * 2 + 1;
  p.ginv = (p.g * (2 - p.g * p.g)) & 31;  // Newton step: g*g == 1 (mod 8) so this is g^-1 (mod 32)
  return p;
}
TL_HD RingParams ring_params(const Geometry& geo, int iter, int ta, int tb) {
  return ring_params_k(iter_key(geo, iter, 3), ta, tb);
}
// B index that lane `a` holds at ring step i.
TL_HD int ring_b(const RingParams& p, int a, int i) { return (a + p.s0 + p.g * i) & 31; }
// Step at which lane a meets B index b.
TL_HD int ring_step(const RingParams& p, int a, int b) { return (p.ginv * (b - a - p.s0)) & 31; }

struct XorParams { int s0, g; };
TL_HD XorParams xor_params_k(uint64_t key /* salt 4 */, int t) {
  const uint64_t h = mix64(key ^ (uint64_t)(uint32_t)t);
  XorParams p;
  p.s0 = (int)(h % 31);
  p.g = (int)((h >> 8) % 30) + 1;  // 1..30, coprime with 31
  return p;
}
TL_HD XorParams xor_params(const Geometry& geo, int iter, int t) { return xor_params_k(iter_key(geo, iter, 4), t); }
// XOR distance used at intra-tile step i (i in [0,31)): a permutation of 1..31.
TL_HD int xor_at(const XorParams& p, int i) { return (p.s0 + p.g * i) % 31 + 1; }

// Sub-round rotation of a cross task.
TL_HD int cross_rot_k(const Geometry& geo, uint64_t key /* salt 5 */, int x, int y) {
  const int lo = x < y ? x : y, hi = x < y ? y : x;
  return (int)(mix64(key ^ (((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi)) % (uint32_t)geo.W);
}
TL_HD int cross_rot(const Geometry& geo, int iter, int x, int y) { return cross_rot_k(geo, iter_key(geo, iter, 5), x, y); }
// Diagonal task: number of tile-level sub-rounds inside one super-block, and the pair a warp
// takes.  Returns false when warp `w` idles in sub-round u.  sb_sel: 0 -> first super-block of
// the task, 1 -> second.  (ia, ib) are tile positions inside that super-block.
TL_HD int diag_subrounds(int W) { return W <= 1 ? 0 : (W + (W & 1)) - 1; }
TL_HD bool diag_pair(int W, int u, int rot, int w, int& sb_sel, int& ia, int& ib) {
  const int Wp = W + (W & 1), Mt = Wp - 1;
  const int uu = (u + rot) % Mt;
  int z;
  if ((W & 1) == 0) {
    const int h = W / 2;
    sb_sel = w / h; z = w % h;
    if (sb_sel > 1) return false;
  } else {
    const int h = (W - 1) / 2;
    if (h == 0 || w >= 2 * h) return false;
    sb_sel = w / h; z = w % h + 1;  // z == 0 would be the bye against the dummy tile
  }
  circle_pair(Wp, uu, z, ia, ib);
  return ia < W && ib < W;
}
TL_HD int diag_rot_k(const Geometry& geo, uint64_t key /* salt 6 */, int q) {
  const int Mt = diag_subrounds(geo.W);
  return Mt > 0 ? (int)(mix64(key ^ (uint64_t)(uint32_t)q) % (uint32_t)Mt) : 0;
}
TL_HD int diag_rot(const Geometry& geo, int iter, int q) { return diag_rot_k(geo, iter_key(geo, iter, 6), q); }

}  // namespace tl
