// oracle/topolow_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's native hot
// path, used as the parity checker by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py.  Nothing under
// topolow_b200/ may import, link or execute this file.
//
// What it restates (citations are relative to /root/reference):
//   * src/optimization.cpp:108-382  optimize_layout_exact_cpp  (the pair loop,
//     cooling, convergence state machine, best-state restore, return values)
//   * src/optimization.cpp:54-81    compute_error_vectorized    (MAE on edges)
//
// Substitutions:
//   1. arma::mat -> plain column-major double buffer, same d*n+i indexing.
//   2. compute_error_vectorized -> scalar loop with the same three masks
//      (:72-75); the final reduction keeps Armadillo's two-accumulator order
//      (accu over a linear proxy) so the sum rounds the same way.
//   3. Rcpp::stop / Rcout / checkUserInterrupt -> status codes / nothing.
//   4. DECLARED DEVIATION: the reference seeds std::mt19937 from
//      std::random_device (:153-154), i.e. it is unseeded.  Here the seed is an
//      argument (order_mode 0) - same std::mt19937 + std::shuffle from the same
//      libstdc++, applied to the same all_pairs vector, so for a given 32-bit
//      seed this draws exactly the permutations the reference would draw if
//      random_device returned that value (pinned by oracle/_ref, which compiles
//      the reference's own source with random_device forced to a constant).
//      order_mode 1 consumes an explicit pair order per iteration instead
//      (used to check structured schedules); order_mode 2 visits a bounded
//      pseudo-random sample of pairs per iteration without materialising the
//      O(N^2) pair array (CPU timing at N = 100k only - the reference cannot
//      run that size at all, src/optimization.cpp:143-150).
//   5. Sparse lookup variant: the dense D[i + j*n] table (:217) is replaced by
//      an open-addressing hash over the edge list; arithmetic is unchanged.
//
// Floating point: build with -ffp-contract=off (see Makefile) so that every
// operation is a separately rounded IEEE double operation, as in an R package
// build for baseline x86-64.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

namespace {

struct PairIdx {  // src/optimization.cpp:100-103
  int i;
  int j;
};

// Target lookup abstraction.  Dense follows :158-159,217,230 literally.
struct DenseLookup {
  const double* dist;
  const int* thresh;
  int64_t n;
  inline void get(int i, int j, double& target, int& type) const {
    const int64_t at = (int64_t)i + (int64_t)j * n;  // column-major [i,j]
    target = dist[at];
    type = thresh ? thresh[at] : 0;
  }
};

// Open addressing on key i*n+j (i<j), value = edge index.  Unmeasured pairs
// return +Inf exactly as the dense matrix would (R/core.R:345).
struct SparseLookup {
  std::vector<int64_t> keys;
  std::vector<int32_t> vals;
  const double* edge_dist;
  const int* edge_thresh;
  int64_t n;
  uint64_t mask;
  static inline uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
  }
  void build(int64_t n_, int64_t n_edges, const int* ei, const int* ej,
             const double* ed, const int* et) {
    n = n_; edge_dist = ed; edge_thresh = et;
    uint64_t cap = 16; while (cap < (uint64_t)n_edges * 2 + 1) cap <<= 1;
    mask = cap - 1; keys.assign(cap, -1); vals.assign(cap, -1);
    for (int64_t e = 0; e < n_edges; ++e) {
      int a = ei[e], b = ej[e]; if (a > b) std::swap(a, b);
      const int64_t key = (int64_t)a * n + b;
      uint64_t h = mix((uint64_t)key) & mask;
      while (keys[h] != -1 && keys[h] != key) h = (h + 1) & mask;
      keys[h] = key; vals[h] = (int32_t)e;  // later duplicates win, like a matrix write
    }
  }
  inline void get(int i, int j, double& target, int& type) const {
    const int64_t key = (int64_t)i * n + j;  // callers pass i<j
    uint64_t h = mix((uint64_t)key) & mask;
    while (true) {
      const int64_t k = keys[h];
      if (k == key) { target = edge_dist[vals[h]]; type = edge_thresh[vals[h]]; return; }
      if (k == -1) { target = std::numeric_limits<double>::infinity(); type = 0; return; }
      h = (h + 1) & mask;
    }
  }
};

// src/optimization.cpp:54-81
inline void edge_error(const double* pos, int64_t n, int dim, int64_t n_edges,
                       const int* ei, const int* ej, const double* target,
                       const int* thresh, double& total, int64_t& count) {
  // Armadillo: accu(abs_errors % contributes) over a linear proxy uses two
  // running sums over even/odd elements, added at the end.
  double acc1 = 0.0, acc2 = 0.0;
  int64_t cnt = 0;
  for (int64_t e = 0; e < n_edges; ++e) {
    const int i = ei[e], j = ej[e];
    double ss = 0.0;
    for (int d = 0; d < dim; ++d) {
      const double delta = pos[(int64_t)d * n + j] - pos[(int64_t)d * n + i];
      ss += delta * delta;
    }
    const double dist = std::sqrt(ss);
    const double abs_err = std::fabs(target[e] - dist);
    const int t = thresh[e];
    const int contributes = (t == 0) + ((t == 1) && (dist < target[e])) +
                            ((t == -1) && (dist > target[e]));
    const double term = abs_err * (double)contributes;
    if ((e & 1) == 0) acc1 += term; else acc2 += term;
    cnt += contributes;
  }
  total = acc1 + acc2;
  count = cnt;
}

// One pair visit: src/optimization.cpp:199-282.
template <class Lookup>
inline void pair_update(double* pos, int64_t n, int dim, int i, int j,
                        const Lookup& lk, const double* deg_plus_one, double k,
                        double c_repulsion) {
  double* pos_i = pos + i;
  double* pos_j = pos + j;
  double dist_sq = 0.0;
  for (int d = 0; d < dim; ++d) {
    const double diff = pos_j[d * n] - pos_i[d * n];
    dist_sq += diff * diff;
  }
  const double dist = std::sqrt(dist_sq);
  const double dist_stable = dist + 0.01;  // :213

  double target_dist; int thresh_type;
  lk.get(i, j, target_dist, thresh_type);
  const bool has_measurement = std::isfinite(target_dist);  // :221
  const double deg_i = deg_plus_one[i];
  const double deg_j = deg_plus_one[j];

  bool apply_spring = false;
  if (has_measurement) {  // :226-243
    if (thresh_type == 0) apply_spring = true;
    else if (thresh_type == 1) apply_spring = (dist < target_dist);
    else apply_spring = (dist > target_dist);
  }
  if (apply_spring) {  // :245-256
    const double factor = 2.0 * k * (target_dist - dist) / dist_stable;
    const double norm_i = 4.0 * deg_i + k;
    const double norm_j = 4.0 * deg_j + k;
    for (int d = 0; d < dim; ++d) {
      const double delta_d = pos_j[d * n] - pos_i[d * n];
      const double force_d = delta_d * factor;
      pos_i[d * n] -= force_d / norm_i;
      pos_j[d * n] += force_d / norm_j;
    }
  } else {  // :257-267 and :269-281 (identical arithmetic)
    const double force_mag = c_repulsion / (2.0 * dist_stable * dist_stable * dist_stable);
    for (int d = 0; d < dim; ++d) {
      const double delta_d = pos_j[d * n] - pos_i[d * n];
      const double force_d = delta_d * force_mag;
      pos_i[d * n] -= force_d / deg_i;
      pos_j[d * n] += force_d / deg_j;
    }
  }
}

// Bijection on [0, 2^bits) used by order_mode 2 (bounded sample); 4-round
// Feistel with cycle walking into [0, P).
inline uint64_t feistel(uint64_t x, int half_bits, uint64_t key) {
  const uint64_t m = (1ULL << half_bits) - 1;
  uint64_t l = x >> half_bits, r = x & m;
  for (int round = 0; round < 4; ++round) {
    const uint64_t f = SparseLookup::mix(r * 0x9e3779b97f4a7c15ULL + key + round) & m;
    const uint64_t nl = r; r = l ^ f; l = nl;
  }
  return (l << half_bits) | r;
}

template <class Lookup>
int run_loop(int64_t n, int dim, const double* initial_positions, const int* degrees,
             int64_t n_edges, const int* edge_i, const int* edge_j, const double* edge_dist,
             const int* edge_thresh, const Lookup& lk, int n_iter, double k0,
             double cooling_rate, double c_repulsion, double relative_epsilon,
             int convergence_window, int convergence_check_freq, int order_mode,
             uint32_t seed, const int32_t* pair_order, int64_t pairs_per_iter,
             double* positions_out, int* converged_out, int* iterations_out,
             double* final_mae_out, double* final_k_out, double* trace_mae, int64_t* visited_out) {
  if (n < 2) return 1;  // "Need at least 2 points for embedding" (:131)

  std::vector<double> pos(initial_positions, initial_positions + n * dim);  // :134
  std::vector<double> deg_plus_one(n);                                       // :137-140
  for (int64_t i = 0; i < n; ++i) deg_plus_one[i] = (double)degrees[i] + 1.0;

  const int64_t num_pairs = n * (n - 1) / 2;  // :143 (int64 here; the reference overflows int at n >= 46342)
  std::vector<PairIdx> all_pairs;
  if (order_mode == 0) {  // :144-150
    all_pairs.reserve(num_pairs);
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) all_pairs.push_back({i, j});
  }
  std::mt19937 rng(seed);  // :153-154 with the declared deviation

  double k = k0;  // :168-179
  double best_mae = std::numeric_limits<double>::max();
  std::vector<double> best_pos = pos;
  double best_k = k0;
  int best_iter = 0;
  int worsening_count = 0;
  const int worsening_patience = convergence_window;
  int converge_count = 0;
  bool converged = false;
  int final_iter = n_iter;
  double final_mae = 0.0;
  int64_t visited = 0;

  if (convergence_check_freq < 1) convergence_check_freq = 10;  // :181

  int half_bits = 1;
  while ((1ULL << (2 * half_bits)) < (uint64_t)num_pairs) ++half_bits;

  for (int iter = 0; iter < n_iter; ++iter) {  // :193
    if (order_mode == 0) {
      std::shuffle(all_pairs.begin(), all_pairs.end(), rng);  // :196
      for (const auto& p : all_pairs)                        // :199
        pair_update(pos.data(), n, dim, p.i, p.j, lk, deg_plus_one.data(), k, c_repulsion);
      visited += num_pairs;
    } else if (order_mode == 1) {
      const int32_t* ord = pair_order + (int64_t)iter * pairs_per_iter * 2;
      for (int64_t p = 0; p < pairs_per_iter; ++p) {
        int i = ord[2 * p], j = ord[2 * p + 1];
        if (i < 0) continue;  // padding slot
        if (i > j) std::swap(i, j);
        pair_update(pos.data(), n, dim, i, j, lk, deg_plus_one.data(), k, c_repulsion);
        ++visited;
      }
    } else {  // bounded sample of a pseudo-random permutation of the pair index space
      const uint64_t key = ((uint64_t)seed << 32) ^ (uint64_t)iter;
      int64_t done = 0;
      for (uint64_t x = 0; done < pairs_per_iter; ++x) {
        uint64_t y = feistel(x, half_bits, key);
        if (y >= (uint64_t)num_pairs) continue;
        // unrank y -> (i<j): row i has (n-1-i) pairs
        const double nn = (double)n - 0.5;
        int64_t i = (int64_t)std::floor(nn - std::sqrt(nn * nn - 2.0 * (double)y));
        while (i > 0 && i * (2 * n - i - 1) / 2 > (int64_t)y) --i;
        while ((i + 1) * (2 * n - i - 2) / 2 <= (int64_t)y) ++i;
        const int64_t j = (int64_t)y - i * (2 * n - i - 1) / 2 + i + 1;
        pair_update(pos.data(), n, dim, (int)i, (int)j, lk, deg_plus_one.data(), k, c_repulsion);
        ++done;
      }
      visited += done;
    }

    k *= (1.0 - cooling_rate);  // :289

    if ((iter + 1) % convergence_check_freq == 0 || iter == n_iter - 1) {  // :294
      double total; int64_t cnt;
      edge_error(pos.data(), n, dim, n_edges, edge_i, edge_j, edge_dist, edge_thresh, total, cnt);
      const double current_error = (cnt > 0) ? total / (double)cnt : 0.0;  // :296
      if (trace_mae) trace_mae[iter] = current_error;

      const double improvement_threshold = best_mae * (1.0 - relative_epsilon);  // :304-305
      const double worsening_threshold = best_mae * (1.0 + relative_epsilon);

      if (current_error < improvement_threshold) {  // :307-314
        best_mae = current_error; best_pos = pos; best_k = k; best_iter = iter + 1;
        worsening_count = 0; converge_count = 0;
      } else if (current_error <= worsening_threshold) {  // :316-338
        if (current_error < best_mae) {
          best_mae = current_error; best_pos = pos; best_k = k; best_iter = iter + 1;
        }
        worsening_count = 0;
        converge_count++;
        if (converge_count >= convergence_window) {
          final_mae = best_mae; final_iter = best_iter; pos = best_pos; k = best_k;
          converged = true;
          break;
        }
      } else {  // :340-356
        converge_count = 0;
        worsening_count++;
        if (worsening_count >= worsening_patience) {
          final_mae = best_mae; final_iter = best_iter; pos = best_pos; k = best_k;
          converged = true;
          break;
        }
      }
    } else if (trace_mae) {
      trace_mae[iter] = std::numeric_limits<double>::quiet_NaN();
    }
    if ((iter + 1) % 10 == 0) {  // :359-361
      bool finite = true;
      for (double v : pos) if (!std::isfinite(v)) { finite = false; break; }
      if (!finite) { *iterations_out = iter + 1; return 2; }
    }
  }

  if (!converged) {  // :368-374
    pos = best_pos; k = best_k; final_mae = best_mae; final_iter = best_iter;
  }
  std::memcpy(positions_out, pos.data(), sizeof(double) * n * dim);  // :375-381
  *converged_out = converged ? 1 : 0;
  *iterations_out = final_iter;
  *final_mae_out = final_mae;
  *final_k_out = k;
  if (visited_out) *visited_out = visited;
  return 0;
}

}  // namespace

extern "C" {

// Status: 0 ok, 1 "Need at least 2 points for embedding", 2 "Numerical
// instability at iteration %d" (iteration in *iterations_out).
//
// dissimilarity_matrix == NULL selects the sparse lookup built from the edge
// list.  order_mode: 0 = std::mt19937(seed) + std::shuffle, 1 = explicit
// pair_order[n_iter][pairs_per_iter][2] (entries with i<0 are skipped),
// 2 = bounded pseudo-random sample of pairs_per_iter pairs per iteration.
int oracle_optimize_layout_exact(
    int64_t n, int dim, const double* initial_positions, const double* dissimilarity_matrix,
    const int* threshold_matrix, const int* degrees, int64_t n_edges, const int* edge_i,
    const int* edge_j, const double* edge_dist, const int* edge_thresh, int n_iter, double k0,
    double cooling_rate, double c_repulsion, double relative_epsilon, int convergence_window,
    int convergence_check_freq, int order_mode, uint32_t seed, const int32_t* pair_order,
    int64_t pairs_per_iter, double* positions_out, int* converged_out, int* iterations_out,
    double* final_mae_out, double* final_k_out, double* trace_mae, int64_t* visited_out) {
  if (dissimilarity_matrix) {
    DenseLookup lk{dissimilarity_matrix, threshold_matrix, n};
    return run_loop(n, dim, initial_positions, degrees, n_edges, edge_i, edge_j, edge_dist,
                    edge_thresh, lk, n_iter, k0, cooling_rate, c_repulsion, relative_epsilon,
                    convergence_window, convergence_check_freq, order_mode, seed, pair_order,
                    pairs_per_iter, positions_out, converged_out, iterations_out, final_mae_out,
                    final_k_out, trace_mae, visited_out);
  }
  SparseLookup lk;
  lk.build(n, n_edges, edge_i, edge_j, edge_dist, edge_thresh);
  return run_loop(n, dim, initial_positions, degrees, n_edges, edge_i, edge_j, edge_dist,
                  edge_thresh, lk, n_iter, k0, cooling_rate, c_repulsion, relative_epsilon,
                  convergence_window, convergence_check_freq, order_mode, seed, pair_order,
                  pairs_per_iter, positions_out, converged_out, iterations_out, final_mae_out,
                  final_k_out, trace_mae, visited_out);
}

// The permutation stream alone: writes the pair order the seeded shuffle draws
// for iterations [0, n_iter) - used to feed replay tests.
int oracle_pair_orders(int64_t n, int n_iter, uint32_t seed, int32_t* out) {
  std::vector<PairIdx> all_pairs;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) all_pairs.push_back({i, j});
  std::mt19937 rng(seed);
  const int64_t P = (int64_t)all_pairs.size();
  for (int it = 0; it < n_iter; ++it) {
    std::shuffle(all_pairs.begin(), all_pairs.end(), rng);
    std::memcpy(out + (int64_t)it * P * 2, all_pairs.data(), sizeof(PairIdx) * P);
  }
  return 0;
}

// MAE helper exposed for unit tests of compute_error_vectorized.
void oracle_edge_error(const double* pos, int64_t n, int dim, int64_t n_edges, const int* ei,
                       const int* ej, const double* target, const int* thresh, double* total,
                       int64_t* count) {
  edge_error(pos, n, dim, n_edges, ei, ej, target, thresh, *total, *count);
}

}  // extern "C"
