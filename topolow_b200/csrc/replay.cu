// topolow_b200/csrc/replay.cu
//
// REPLAY MODE (TOPOLOW_MODE_REPLAY): executes a given sequential pair order -
// the std::mt19937 + std::shuffle stream of src/optimization.cpp:153-154,196 for a
// chosen seed, or an explicit order - with exactly the sequential semantics of
// the reference's Gauss-Seidel pair loop (src/optimization.cpp:199-282).
//
// How: the host cuts each iteration's order into dependency levels
// (level(p) = 1 + max(level of the previous pair touching i, ... touching j)).
// Pairs of one level are point-disjoint, so their updates commute; levels run
// in order with a CTA barrier between them.  One CTA per fit, positions staged
// in shared memory, FP64 with one IEEE rounding per reference operation
// (__dadd_rn/__dmul_rn/__ddiv_rn/__dsqrt_rn: no FMA contraction), so the result
// is bit-comparable with the CPU loop.  The edge MAE (src/optimization.cpp:54-81)
// is summed in Armadillo's two-accumulator order for the same reason.
#include "replay.h"

#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

namespace tl {

namespace {

struct ReplayDev {
  double* pos;        // [n][dim] row-major working positions
  double* best_pos;   // [n][dim]
  const double* dp1;  // deg + 1
  const int* ei; const int* ej; const double* et; const int* ety; // COO edges for the MAE
  double* terms;      // [E] scratch
  int* contrib;       // [E] scratch
  FitState* state;
  double* trace;      // [n_iter] or null
  int n, dim, n_edges;
};

TL_D void pair_update_exact(double* __restrict__ pi, double* __restrict__ pj, int dim, double target,
                            int type, double deg_i, double deg_j, double k, double c_rep) {
  // src/optimization.cpp:207-213
  double dist_sq = 0.0;
  for (int d = 0; d < dim; ++d) {
    const double diff = __dsub_rn(pj[d], pi[d]);
    dist_sq = __dadd_rn(dist_sq, __dmul_rn(diff, diff));
  }
  const double dist = __dsqrt_rn(dist_sq);
  const double dist_stable = __dadd_rn(dist, 0.01);
  bool spring = false;
  if (isfinite(target)) {  // :221-243
    spring = (type == 0) ? true : (type == 1 ? (dist < target) : (dist > target));
  }
  if (spring) {  // :247-256
    const double factor = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, k), __dsub_rn(target, dist)), dist_stable);
    const double norm_i = __dadd_rn(__dmul_rn(4.0, deg_i), k);
    const double norm_j = __dadd_rn(__dmul_rn(4.0, deg_j), k);
    for (int d = 0; d < dim; ++d) {
      const double delta = __dsub_rn(pj[d], pi[d]);
      const double force = __dmul_rn(delta, factor);
      pi[d] = __dsub_rn(pi[d], __ddiv_rn(force, norm_i));
      pj[d] = __dadd_rn(pj[d], __ddiv_rn(force, norm_j));
    }
  } else {  // :259-266, :273-280
    const double ds3 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, dist_stable), dist_stable), dist_stable);
    const double force_mag = __ddiv_rn(c_rep, ds3);
    for (int d = 0; d < dim; ++d) {
      const double delta = __dsub_rn(pj[d], pi[d]);
      const double force = __dmul_rn(delta, force_mag);
      pi[d] = __dsub_rn(pi[d], __ddiv_rn(force, deg_i));
      pj[d] = __dadd_rn(pj[d], __ddiv_rn(force, deg_j));
    }
  }
}

// A pair visit as the host ships it: the points and what the dense matrices of the reference call
// hold for them (target = +Inf when the pair is unmeasured; src/optimization.cpp:217,230).
struct __align__(16) PairRec {
  double target;
  uint32_t ij;     // i << 16 | j
  int32_t type;
};

constexpr int kLvlCap = 4096;   // level offsets of one iteration staged in shared memory

// One launch = `n_it` iterations of one fit.  pairs: records sorted by level within each iteration;
// lvl_off: per-iteration level offsets into pairs (absolute), terminated; it_lvl: [n_it+1] index of
// each iteration's first entry in lvl_off.
__global__ void __launch_bounds__(1024)
replay_kernel(ReplayDev dv, FitParams prm, const PairRec* __restrict__ pairs,
              const int* __restrict__ lvl_off, const int* __restrict__ it_lvl, int n_it,
              int use_smem, volatile int* host_flag) {
  extern __shared__ double smem_pos[];
  __shared__ int s_lvl[kLvlCap];
  __shared__ FitState st;
  __shared__ double red_a, red_b;
  __shared__ long long red_cnt;
  __shared__ int red_bad;

  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = dv.n, dim = dv.dim;
  double* P = use_smem ? smem_pos : dv.pos;
  if (tid == 0) st = *dv.state;
  if (use_smem)
    for (int x = tid; x < n * dim; x += nt) P[x] = dv.pos[x];
  __syncthreads();

  for (int t = 0; t < n_it; ++t) {
    if (st.stop) break;  // uniform (shared)
    const int iter = st.iter;
    const double k = st.k;
    const int l0 = it_lvl[t], l1 = it_lvl[t + 1] - 1;  // levels are [lvl_off[l], lvl_off[l+1])
    const int nl = l1 - l0;
    const bool staged = nl + 1 <= kLvlCap;
    if (staged) {
      for (int x = tid; x <= nl; x += nt) s_lvl[x] = lvl_off[l0 + x];
      __syncthreads();
    }
    // the record of the next level is fetched while the current one is computed
    PairRec nxt;
    int nb = staged ? s_lvl[0] : lvl_off[l0], ne = nl > 0 ? (staged ? s_lvl[1] : lvl_off[l0 + 1]) : nb;
    if (nb + tid < ne) nxt = pairs[nb + tid];
    for (int l = 0; l < nl; ++l) {
      const int b = nb, e = ne;
      const PairRec cur = nxt;
      if (l + 1 < nl) {
        nb = e;
        ne = staged ? s_lvl[l + 2] : lvl_off[l0 + l + 2];
        if (nb + tid < ne) nxt = pairs[nb + tid];
      }
      for (int p = b + tid; p < e; p += nt) {
        const PairRec r = (p == b + tid) ? cur : pairs[p];
        const int i = r.ij >> 16, j = r.ij & 0xffff;
        pair_update_exact(P + (size_t)i * dim, P + (size_t)j * dim, dim, r.target, r.type, dv.dp1[i], dv.dp1[j], k,
                          prm.c_repulsion);
      }
      __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
      st.k = __dmul_rn(st.k, __dsub_rn(1.0, prm.cooling_rate));  // :289
      st.pair_updates += (unsigned long long)(lvl_off[l1] - lvl_off[l0]);
    }
    const bool check = is_check_iter(iter, prm);
    if (check) {  // :294-357
      for (int e = tid; e < dv.n_edges; e += nt) {
        const double* a = P + (size_t)dv.ei[e] * dim;
        const double* bq = P + (size_t)dv.ej[e] * dim;
        double ss = 0.0;
        for (int d = 0; d < dim; ++d) {
          const double df = __dsub_rn(bq[d], a[d]);
          ss = __dadd_rn(ss, __dmul_rn(df, df));
        }
        const double dist = __dsqrt_rn(ss);
        const double tg = dv.et[e];
        const int ty = dv.ety[e];
        const int c = (ty == 0) + ((ty == 1) && (dist < tg)) + ((ty == -1) && (dist > tg));
        dv.terms[e] = __dmul_rn(fabs(__dsub_rn(tg, dist)), (double)c);
        dv.contrib[e] = c;
      }
      __syncthreads();
      // Armadillo accu(): even elements into one running sum, odd into another.
      const int t_odd = nt > 32 ? 32 : 0, t_cnt = nt > 64 ? 64 : 0;
      if (tid == 0) { double a = 0.0; for (int e = 0; e < dv.n_edges; e += 2) a = __dadd_rn(a, dv.terms[e]); red_a = a; }
      if (tid == t_odd) { double a = 0.0; for (int e = 1; e < dv.n_edges; e += 2) a = __dadd_rn(a, dv.terms[e]); red_b = a; }
      if (tid == t_cnt) { long long c = 0; for (int e = 0; e < dv.n_edges; ++e) c += dv.contrib[e]; red_cnt = c; }
      __syncthreads();
      if (tid == 0) {
        controller_check(st, prm, iter, __dadd_rn(red_a, red_b), red_cnt);
        if (dv.trace) dv.trace[iter] = st.last_error;
      }
      __syncthreads();
      if (st.snapshot) {
        for (int x = tid; x < n * dim; x += nt) dv.best_pos[x] = P[x];
      }
    }
    if (!st.stop && (iter + 1) % 10 == 0) {  // :359-361
      if (tid == 0) red_bad = 0;
      __syncthreads();
      int bad = 0;
      for (int x = tid; x < n * dim; x += nt) bad |= !isfinite(P[x]);
      if (bad) red_bad = 1;
      __syncthreads();
      if (tid == 0 && red_bad) { st.status = 2; st.fail_iter = iter + 1; st.stop = 1; }
    }
    if (tid == 0) st.iter = iter + 1;
    __syncthreads();
  }
  if (use_smem)
    for (int x = tid; x < n * dim; x += nt) dv.pos[x] = P[x];
  if (tid == 0) {
    *dv.state = st;
    if (host_flag) { host_flag[1] = st.iter; __threadfence_system(); host_flag[0] = st.stop; }
  }
}

}  // namespace

int replay_max_n() { return 8192; }

void run_replay(const topolow_problem& pb, const topolow_params& pr, topolow_result& res,
                topolow_interrupt_fn poll, void* user) {
  const int n = (int)pb.n, dim = pb.ndim;
  const int64_t E = pb.n_edges;
  const int64_t P = (int64_t)n * (n - 1) / 2;
  FitParams prm{pr.n_iter, pr.k0, pr.cooling_rate, pr.c_repulsion, pr.relative_epsilon,
                pr.convergence_window, pr.convergence_check_freq};

  // ---- host-side images -------------------------------------------------
  std::vector<double> h_pos((size_t)n * dim);
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) h_pos[(size_t)i * dim + d] = pb.initial_positions[(size_t)d * n + i];
  std::vector<double> h_dist((size_t)n * n, INFINITY);
  std::vector<int8_t> h_thr((size_t)n * n, 0);
  std::vector<int> h_ei(E), h_ej(E);
  for (int64_t e = 0; e < E; ++e) {
    int a = pb.edge_i[e], b = pb.edge_j[e];
    if (a < 0 || b < 0 || a >= n || b >= n || a == b) throw std::invalid_argument("edge index out of range");
    h_ei[e] = a; h_ej[e] = b;
    if (a > b) std::swap(a, b);
    h_dist[(size_t)a * n + b] = pb.edge_dist[e];
    h_thr[(size_t)a * n + b] = (int8_t)pb.edge_thresh[e];
  }
  std::vector<double> h_dp1(n);
  for (int i = 0; i < n; ++i) h_dp1[i] = (double)pb.degrees[i] + 1.0;

  ReplayDev dv{};
  dv.n = n; dv.dim = dim; dv.n_edges = (int)E;
  DeviceBuf<double> d_pos((size_t)n * dim), d_best((size_t)n * dim), d_dp1(n), d_et(E), d_terms(E);
  DeviceBuf<int> d_ei(E), d_ej(E), d_ety(E), d_contrib(E);
  DeviceBuf<FitState> d_state(1);
  DeviceBuf<double> d_trace(res.trace_mae ? pr.n_iter : 0);
  TL_CUDA(cudaMemcpy(d_pos, h_pos.data(), sizeof(double) * n * dim, cudaMemcpyHostToDevice));
  TL_CUDA(cudaMemcpy(d_best, h_pos.data(), sizeof(double) * n * dim, cudaMemcpyHostToDevice));
  TL_CUDA(cudaMemcpy(d_dp1, h_dp1.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  if (E) {
    TL_CUDA(cudaMemcpy(d_ei, h_ei.data(), sizeof(int) * E, cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy(d_ej, h_ej.data(), sizeof(int) * E, cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy(d_et, pb.edge_dist, sizeof(double) * E, cudaMemcpyHostToDevice));
    TL_CUDA(cudaMemcpy(d_ety, pb.edge_thresh, sizeof(int) * E, cudaMemcpyHostToDevice));
  }
  if (res.trace_mae) {
    std::vector<double> nanv(pr.n_iter, NAN);
    TL_CUDA(cudaMemcpy(d_trace, nanv.data(), sizeof(double) * pr.n_iter, cudaMemcpyHostToDevice));
  }
  FitState h_state; state_init(h_state, prm);
  TL_CUDA(cudaMemcpy(d_state, &h_state, sizeof h_state, cudaMemcpyHostToDevice));
  dv.pos = d_pos; dv.best_pos = d_best; dv.dp1 = d_dp1;
  dv.ei = d_ei; dv.ej = d_ej; dv.et = d_et; dv.ety = d_ety; dv.terms = d_terms; dv.contrib = d_contrib;
  dv.state = d_state; dv.trace = res.trace_mae ? (double*)d_trace : nullptr;

  // ---- launch geometry ----------------------------------------------------
  const size_t smem_need = sizeof(double) * (size_t)n * dim;
  const int use_smem = smem_need <= 160 * 1024;
  if (use_smem && smem_need > 48 * 1024)
    TL_CUDA(cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
  int threads = ((n / 4 + 31) / 32) * 32;
  threads = std::min(1024, std::max(64, threads));

  // ---- chunked, double-buffered order stream --------------------------------
  const int64_t chunk_pairs_target = 1 << 21;
  const int chunk_iters = (int)std::max<int64_t>(1, std::min<int64_t>(chunk_pairs_target / std::max<int64_t>(P, 1), 64));
  const int64_t ppi = pr.pair_order ? pr.pairs_per_iter : P;
  const size_t cap_pairs = (size_t)chunk_iters * ppi;
  const size_t cap_lvls = (size_t)chunk_iters * (ppi + 2);
  PinnedBuf<PairRec> h_pairs[2] = {PinnedBuf<PairRec>(cap_pairs), PinnedBuf<PairRec>(cap_pairs)};
  PinnedBuf<int> h_lvl[2] = {PinnedBuf<int>(cap_lvls), PinnedBuf<int>(cap_lvls)};
  PinnedBuf<int> h_itl[2] = {PinnedBuf<int>(chunk_iters + 1), PinnedBuf<int>(chunk_iters + 1)};
  DeviceBuf<PairRec> d_pairs[2] = {DeviceBuf<PairRec>(cap_pairs), DeviceBuf<PairRec>(cap_pairs)};
  DeviceBuf<int> d_lvl[2] = {DeviceBuf<int>(cap_lvls), DeviceBuf<int>(cap_lvls)};
  DeviceBuf<int> d_itl[2] = {DeviceBuf<int>(chunk_iters + 1), DeviceBuf<int>(chunk_iters + 1)};
  EventGuard copied[2] = {EventGuard(cudaEventDisableTiming), EventGuard(cudaEventDisableTiming)};
  PinnedBuf<int> h_flag_buf(2, cudaHostAllocMapped);
  volatile int* h_flag = h_flag_buf.p;
  int* d_flag = nullptr;
  h_flag[0] = 0; h_flag[1] = 0;
  TL_CUDA(cudaHostGetDevicePointer((void**)&d_flag, (void*)h_flag, 0));
  StreamGuard stream;
  EventGuard ev0, ev1;
  TL_CUDA(cudaEventRecord(ev0, stream));

  struct PairIdx { int i, j; };
  std::vector<PairIdx> all_pairs;
  if (!pr.pair_order) {  // src/optimization.cpp:143-150
    all_pairs.reserve(P);
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) all_pairs.push_back({i, j});
  }
  std::mt19937 rng((uint32_t)pr.seed);  // :153-154 with an explicit seed
  std::vector<int> last(n), lvl_of, counts;
  lvl_of.resize(ppi);

  bool interrupted = false;
  int launched_iters = 0;
  for (int c = 0; launched_iters < pr.n_iter; ++c) {
    if (h_flag[0]) break;  // device reported stop
    if (poll && poll(user)) { interrupted = true; break; }
    const int b = c & 1;
    if (c >= 2) TL_CUDA(cudaEventSynchronize(copied[b]));
    const int nit = std::min(chunk_iters, pr.n_iter - launched_iters);
    size_t np = 0, nl = 0;
    for (int t = 0; t < nit; ++t) {
      const int it = launched_iters + t;
      const PairIdx* ord; int64_t cnt;
      std::vector<PairIdx> tmp;
      if (pr.pair_order) {
        const int32_t* src = pr.pair_order + (size_t)it * ppi * 2;
        tmp.reserve(ppi);
        for (int64_t p = 0; p < ppi; ++p) {
          int i = src[2 * p], j = src[2 * p + 1];
          if (i < 0) continue;
          if (j < 0 || i >= n || j >= n || i == j) throw std::invalid_argument("pair_order entry out of range");
          if (i > j) std::swap(i, j);
          tmp.push_back({i, j});
        }
        ord = tmp.data(); cnt = (int64_t)tmp.size();
      } else {
        std::shuffle(all_pairs.begin(), all_pairs.end(), rng);  // :196
        ord = all_pairs.data(); cnt = P;
      }
      // dependency levels
      std::fill(last.begin(), last.end(), 0);
      int max_l = 0;
      for (int64_t p = 0; p < cnt; ++p) {
        const int l = std::max(last[ord[p].i], last[ord[p].j]) + 1;
        last[ord[p].i] = l; last[ord[p].j] = l; lvl_of[p] = l;
        if (l > max_l) max_l = l;
      }
      counts.assign(max_l + 2, 0);
      for (int64_t p = 0; p < cnt; ++p) counts[lvl_of[p]]++;
      // offsets (absolute in the chunk's pair buffer)
      h_itl[b][t] = (int)nl;
      int run = (int)np;
      for (int l = 1; l <= max_l; ++l) { h_lvl[b][nl++] = run; const int cl = counts[l]; counts[l] = run; run += cl; }
      h_lvl[b][nl++] = run;
      for (int64_t p = 0; p < cnt; ++p) {
        const size_t at = (size_t)ord[p].i * n + ord[p].j;
        PairRec& r = h_pairs[b][counts[lvl_of[p]]++];
        r.target = h_dist[at];
        r.ij = ((uint32_t)ord[p].i << 16) | (uint32_t)ord[p].j;
        r.type = h_thr[at];
      }
      np = run;
    }
    h_itl[b][nit] = (int)nl;
    TL_CUDA(cudaMemcpyAsync(d_pairs[b], h_pairs[b], sizeof(PairRec) * np, cudaMemcpyHostToDevice, stream));
    TL_CUDA(cudaMemcpyAsync(d_lvl[b], h_lvl[b], sizeof(int) * nl, cudaMemcpyHostToDevice, stream));
    TL_CUDA(cudaMemcpyAsync(d_itl[b], h_itl[b], sizeof(int) * (nit + 1), cudaMemcpyHostToDevice, stream));
    TL_CUDA(cudaEventRecord(copied[b], stream));
    replay_kernel<<<1, threads, use_smem ? smem_need : 0, stream>>>(dv, prm, d_pairs[b], d_lvl[b], d_itl[b],
                                                                   nit, use_smem, d_flag);
    TL_CUDA(cudaGetLastError());
    launched_iters += nit;
  }
  TL_CUDA(cudaEventRecord(ev1, stream));
  TL_CUDA(cudaStreamSynchronize(stream));
  float ms = 0.f; TL_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));

  // ---- results: always the best snapshot (src/optimization.cpp:368-381) ---------------
  TL_CUDA(cudaMemcpy(&h_state, d_state, sizeof h_state, cudaMemcpyDeviceToHost));
  TL_CUDA(cudaMemcpy(h_pos.data(), d_best, sizeof(double) * n * dim, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) res.positions[(size_t)d * n + i] = h_pos[(size_t)i * dim + d];
  res.converged = h_state.converged;
  res.iterations = h_state.best_iter;
  res.final_mae = h_state.best_mae;
  res.final_k = h_state.best_k;
  res.iterations_run = h_state.iter;
  res.pair_updates = (int64_t)h_state.pair_updates;
  res.device_ms = ms;
  res.status = interrupted ? TOPOLOW_ERR_INTERRUPTED : (h_state.status == 2 ? TOPOLOW_ERR_NONFINITE : TOPOLOW_OK);
  res.fail_iter = h_state.fail_iter;
  if (res.trace_mae) TL_CUDA(cudaMemcpy(res.trace_mae, d_trace, sizeof(double) * pr.n_iter, cudaMemcpyDeviceToHost));

}

}  // namespace tl
