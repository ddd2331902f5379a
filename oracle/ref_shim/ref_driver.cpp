// oracle/ref_shim/ref_driver.cpp
//
// TEST INFRASTRUCTURE ONLY.  Compiles the reference's own native translation
// unit, unmodified and from where it lies (REFERENCE_SRC is set by
// oracle/Makefile to /root/reference/src/optimization.cpp), against the
// stand-in headers in this directory, and wraps its one exported function
// (src/optimization.cpp:108-126, the signature behind
// _topolow_optimize_layout_exact_cpp, src/RcppExports.cpp:16-39) in a C ABI.
// Output goes to oracle/_ref/ only.  No reference source is copied.
#include "RcppArmadillo.h"

namespace topolow_ref_shim {
unsigned g_seed = 0;
}

#include REFERENCE_SRC

extern "C" int ref_optimize_layout_exact(
    int n, int dim, double* initial_positions, double* dissimilarity_matrix, int* threshold_matrix,
    const int* degrees, int n_edges, const int* edge_i, const int* edge_j, const double* edge_dist,
    const int* edge_thresh, int n_iter, double k0, double cooling_rate, double c_repulsion,
    double relative_epsilon, int convergence_window, int convergence_check_freq, unsigned seed,
    double* positions_out, int* converged_out, int* iterations_out, double* final_mae_out,
    double* final_k_out, char* errbuf, int errbuf_len) {
  topolow_ref_shim::g_seed = seed;
  try {
    Rcpp::List res = optimize_layout_exact_cpp(
        Rcpp::NumericMatrix(initial_positions, n, dim), Rcpp::NumericMatrix(dissimilarity_matrix, n, n),
        Rcpp::IntegerMatrix(threshold_matrix, n, n), Rcpp::IntegerVector(degrees, n),
        Rcpp::IntegerVector(edge_i, n_edges), Rcpp::IntegerVector(edge_j, n_edges),
        Rcpp::NumericVector(edge_dist, n_edges), Rcpp::IntegerVector(edge_thresh, n_edges), n_iter, k0,
        cooling_rate, c_repulsion, relative_epsilon, convergence_window, convergence_check_freq, false);
    const arma::mat& p = res.items["positions"].m;
    std::copy(p.m.begin(), p.m.end(), positions_out);
    *converged_out = (int)res.items["converged"].d;
    *iterations_out = (int)res.items["iterations"].d;
    *final_mae_out = res.items["final_mae"].d;
    *final_k_out = res.items["final_k"].d;
    return 0;
  } catch (const Rcpp::stop_error& e) {
    if (errbuf && errbuf_len > 0) std::snprintf(errbuf, errbuf_len, "%s", e.what());
    return 1;
  }
}
