// oracle/relaxed_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement (FP64) of the ROW-BLOCK ("relaxed", owner-computes)
// iteration that topolow_b200/csrc/rowblock.cu runs on the GPU - the partitioning SURVEY.md section 8e
// describes for one large map: every point's update is computed by the owner of its row against a copy
// of all other points.  Used by tests/ to check the CUDA kernels (same arithmetic, FP64 here) and to
// compare the scheme statistically with the reference's sequential loop (oracle/topolow_oracle.cpp).
// Nothing under topolow_b200/ may import, link or execute this file.
//
// One iteration on the snapshot P of all positions (citations: /root/reference/src/optimization.cpp):
//   1. repulsion, Jacobi over all ordered pairs (:269-281 applied from the snapshot):
//        R_i = - sum_{j != i} (P_j - P_i) * c / (2 (|P_j - P_i| + 0.01)^3) / (deg_i + 1)
//   2. springs, Gauss-Seidel along the point's own measured pairs (:226-256), partners at the snapshot:
//        x = P_i;  for every record (j, target, type) of row i (partners ascending by slot; the rows
//        of a slice of 32 slots share a start offset drawn per iteration and walk the slice's width cyclically):
//          delta = P_j - x, dist = |delta|, ds = dist + 0.01
//          spring iff type == 0, or '>' and dist < target, or '<' and dist > target   (:237-243)
//          if spring:  x -= delta * (2 k (target - dist) / ds) / (4 (deg_i + 1) + k)   (:246-253, own endpoint)
//                      x += (P_j - P_i) * c / (2 (|P_j - P_i| + 0.01)^3) / (deg_i + 1)   (takes back what
//                           step 1 applied to this pair: a pair in spring state gets no repulsion, :226-256)
//          else: nothing (a satisfied threshold is repelled like an unmeasured pair, :257-267: step 1 did it)
//      P'_i = x + R_i   (steps 1 and 2 both read only the snapshot: the GPU runs them side by side)
//   3. k *= 1 - cooling_rate (:289); every check_freq iterations and on the last one the edge MAE
//      (:54-81) on P' and the three-way controller with best-state snapshot (:303-357,368-374).
// Every unordered pair is visited once from each side per iteration; each side moves only its own
// endpoint by exactly the amount the reference's pair visit would move it.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rec { int32_t j; int32_t type; double target; };

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// body(i) for i in [0, n) on all host cores; iterations are independent.
template <class F>
void parallel_for(int64_t n, F body) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if ((int64_t)nt > n) nt = (unsigned)(n > 0 ? n : 1);
  if (nt == 1) { for (int64_t i = 0; i < n; ++i) body(i); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; ++t)
    pool.emplace_back([=] { for (int64_t i = t; i < n; i += nt) body(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// positions: column-major n x dim (like an R matrix) in and out.  slot_order (optional, length n):
// slot_order[s] = point held by slot s - the order rows are grouped and (within a row) partners are
// visited in; NULL = identity.  rotate != 0: the walk over the rows of slice s / 32 starts at
// mix64(seed ^ mix64(iter << 32 | slice)) % width(slice).  Returns 0, or 2 on non-finite positions
// (fail_iter in *iterations).
int relaxed_optimize_layout(int64_t n, int dim, const double* init, const int* degrees, int64_t n_edges,
                            const int* edge_i, const int* edge_j, const double* edge_dist, const int* edge_thresh,
                            int n_iter, double k0, double cooling_rate, double c_repulsion, double relative_epsilon,
                            int convergence_window, int convergence_check_freq, const int32_t* slot_order,
                            int rotate, uint64_t seed,
                            double* out_positions, int* out_converged, int* out_iterations, double* out_final_mae,
                            double* out_final_k, double* trace, int* out_iterations_run) {
  if (n < 2) return 1;
  if (convergence_check_freq < 1) convergence_check_freq = 10;
  std::vector<int32_t> point_of_slot(n), slot_of_point(n);
  for (int64_t s = 0; s < n; ++s) point_of_slot[s] = slot_order ? slot_order[s] : (int32_t)s;
  for (int64_t s = 0; s < n; ++s) slot_of_point[point_of_slot[s]] = (int32_t)s;
  // CSR by slot, partners ascending by slot
  std::vector<int64_t> off(n + 1, 0);
  for (int64_t e = 0; e < n_edges; ++e) { off[slot_of_point[edge_i[e]] + 1]++; off[slot_of_point[edge_j[e]] + 1]++; }
  for (int64_t s = 0; s < n; ++s) off[s + 1] += off[s];
  std::vector<Rec> recs(2 * n_edges);
  {
    std::vector<int64_t> cur(off.begin(), off.end() - 1);
    for (int64_t e = 0; e < n_edges; ++e) {
      const int32_t a = slot_of_point[edge_i[e]], b = slot_of_point[edge_j[e]];
      recs[cur[a]++] = Rec{b, edge_thresh[e], edge_dist[e]};
      recs[cur[b]++] = Rec{a, edge_thresh[e], edge_dist[e]};
    }
    for (int64_t s = 0; s < n; ++s)
      std::sort(recs.begin() + off[s], recs.begin() + off[s + 1], [](const Rec& x, const Rec& y) {
        if (x.j != y.j) return x.j < y.j;
        const int tx = x.type == 0 ? 0 : (x.type == 1 ? 1 : 2), ty = y.type == 0 ? 0 : (y.type == 1 ? 1 : 2);
        if (tx != ty) return tx < ty;
        return x.target < y.target;
      });
  }
  std::vector<int64_t> slice_width((n + 31) / 32, 0);
  for (int64_t s = 0; s < n; ++s) slice_width[s / 32] = std::max(slice_width[s / 32], off[s + 1] - off[s]);
  // slot-major AoS positions
  std::vector<double> P((size_t)n * dim), Pn((size_t)n * dim), best((size_t)n * dim), dp1(n);
  for (int64_t s = 0; s < n; ++s) {
    const int64_t i = point_of_slot[s];
    for (int d = 0; d < dim; ++d) P[s * dim + d] = init[(size_t)d * n + i];
    dp1[s] = (double)degrees[i] + 1.0;
  }
  best = P;
  double k = k0, best_mae = DBL_MAX, best_k = k0;
  int best_iter = 0, worsening = 0, conv_count = 0, converged = 0, iters_run = 0;
  const double c_half = 0.5 * c_repulsion;
  std::vector<double> R((size_t)n * dim);
  int status = 0;

  for (int iter = 0; iter < n_iter; ++iter) {
    // ---- 1. repulsion from the snapshot ----
    parallel_for(n, [&](int64_t i) {
      double acc[64];
      for (int d = 0; d < dim; ++d) acc[d] = 0.0;
      const double* pi = &P[i * dim];
      for (int64_t j = 0; j < n; ++j) {
        const double* pj = &P[j * dim];
        double d2 = 0.0, delta[64];
        for (int d = 0; d < dim; ++d) { delta[d] = pj[d] - pi[d]; d2 += delta[d] * delta[d]; }
        const double ds = std::sqrt(d2) + 0.01;
        const double w = 1.0 / (ds * ds * ds);
        for (int d = 0; d < dim; ++d) acc[d] += delta[d] * w;   // j == i contributes delta = 0
      }
      for (int d = 0; d < dim; ++d) R[i * dim + d] = -acc[d] * c_half / dp1[i];
    });
    // ---- 2. springs ----
    parallel_for(n, [&](int64_t i) {
      double x[64];
      const double* pi = &P[i * dim];
      for (int d = 0; d < dim; ++d) x[d] = pi[d];
      const int64_t len = off[i + 1] - off[i];
      const int64_t width = slice_width[i / 32];
      const int64_t start = (rotate && width > 0)
          ? (int64_t)(mix64(seed ^ mix64(((uint64_t)(uint32_t)iter << 32) | (uint64_t)(i / 32))) % (uint64_t)width) : 0;
      const double rnorm = 1.0 / (4.0 * dp1[i] + k), rdeg = c_half / dp1[i];
      for (int64_t t = 0; t < width; ++t) {
        int64_t at = start + t; if (at >= width) at -= width;
        if (at >= len) continue;                       // padding of the slice
        const Rec& r = recs[off[i] + at];
        const double* pj = &P[(size_t)r.j * dim];
        double d2 = 0.0, delta[64];
        for (int d = 0; d < dim; ++d) { delta[d] = pj[d] - x[d]; d2 += delta[d] * delta[d]; }
        const double dist = std::sqrt(d2), ds = dist + 0.01;
        const bool spring = r.type == 0 || (r.type > 0 ? dist < r.target : dist > r.target);
        if (!spring) continue;
        const double f = 2.0 * k * (r.target - dist) / ds * rnorm;
        double e2 = 0.0, d0[64];
        for (int d = 0; d < dim; ++d) { d0[d] = pj[d] - pi[d]; e2 += d0[d] * d0[d]; }
        const double ds0 = std::sqrt(e2) + 0.01;
        const double w0 = rdeg / (ds0 * ds0 * ds0);
        for (int d = 0; d < dim; ++d) x[d] += -delta[d] * f + d0[d] * w0;
      }
      for (int d = 0; d < dim; ++d) Pn[i * dim + d] = x[d] + R[i * dim + d];
    });
    P.swap(Pn);
    k *= (1.0 - cooling_rate);
    iters_run = iter + 1;
    // ---- 3. checks ----
    const bool check = ((iter + 1) % convergence_check_freq == 0) || (iter == n_iter - 1);
    if (check) {
      double tot = 0.0; int64_t cnt = 0;
      for (int64_t i = 0; i < n; ++i)
        for (int64_t t = off[i]; t < off[i + 1]; ++t) {
          const Rec& r = recs[t];
          if (r.j <= i) continue;
          double d2 = 0.0;
          for (int d = 0; d < dim; ++d) { const double df = P[(size_t)r.j * dim + d] - P[i * dim + d]; d2 += df * df; }
          const double dist = std::sqrt(d2);
          const bool contributes = r.type == 0 || (r.type > 0 ? dist < r.target : dist > r.target);
          if (contributes) { tot += std::fabs(r.target - dist); ++cnt; }
        }
      const double err = cnt > 0 ? tot / (double)cnt : 0.0;
      if (trace) trace[iter] = err;
      const double imp = best_mae * (1.0 - relative_epsilon), wor = best_mae * (1.0 + relative_epsilon);
      bool stop = false;
      if (err < imp) { best_mae = err; best_k = k; best_iter = iter + 1; best = P; worsening = 0; conv_count = 0; }
      else if (err <= wor) {
        if (err < best_mae) { best_mae = err; best_k = k; best_iter = iter + 1; best = P; }
        worsening = 0;
        if (++conv_count >= convergence_window) { converged = 1; stop = true; }
      } else {
        conv_count = 0;
        if (++worsening >= convergence_window) { converged = 1; stop = true; }
      }
      if (stop) break;
    }
    if ((iter + 1) % 10 == 0) {
      bool bad = false;
      for (size_t x = 0; x < P.size(); ++x) if (!std::isfinite(P[x])) { bad = true; break; }
      if (bad) { status = 2; best_iter = iter + 1; break; }
    }
  }
  for (int64_t s = 0; s < n; ++s) {
    const int64_t i = point_of_slot[s];
    for (int d = 0; d < dim; ++d) out_positions[(size_t)d * n + i] = best[s * dim + d];
  }
  *out_converged = converged; *out_iterations = best_iter; *out_final_mae = best_mae; *out_final_k = best_k;
  if (out_iterations_run) *out_iterations_run = iters_run;
  return status;
}

}  // extern "C"
