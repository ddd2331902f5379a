/* integration/r_shim.c - the .Call wrapper around libtopolow_b200.so.
 * Ships as source: no R headers exist in the build image, so it is neither compiled nor tested here. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "topolow_b200.h"

static int poll_interrupt(void* u) { (void)u; R_CheckUserInterrupt(); return 0; }

/* Same 16 SEXPs, same order, as _topolow_optimize_layout_exact_cpp (src/RcppExports.cpp:16). */
SEXP _topolow_optimize_layout_b200(SEXP init, SEXP dmat, SEXP tmat, SEXP degrees, SEXP edge_i, SEXP edge_j,
                                   SEXP edge_dist, SEXP edge_thresh, SEXP n_iter, SEXP k0, SEXP cooling,
                                   SEXP c_rep, SEXP rel_eps, SEXP window, SEXP freq, SEXP verbose) {
  const int n = Rf_nrows(init), d = Rf_ncols(init);
  topolow_problem pb = { n, d, XLENGTH(edge_i), INTEGER(edge_i), INTEGER(edge_j), REAL(edge_dist),
                         INTEGER(edge_thresh), INTEGER(degrees), REAL(init) };   /* inputs are read only */
  topolow_params pr = {0};
  pr.n_iter = Rf_asInteger(n_iter); pr.k0 = Rf_asReal(k0); pr.cooling_rate = Rf_asReal(cooling);
  pr.c_repulsion = Rf_asReal(c_rep); pr.relative_epsilon = Rf_asReal(rel_eps);
  pr.convergence_window = Rf_asInteger(window); pr.convergence_check_freq = Rf_asInteger(freq);
  pr.verbose = Rf_asLogical(verbose);
  pr.seed = (uint64_t)(unif_rand() * 4294967296.0);   /* inside GetRNGstate()/PutRNGstate(): set.seed() now fixes the run */
  SEXP pos = PROTECT(Rf_allocMatrix(REALSXP, n, d));
  topolow_result rs = {0};
  rs.positions = REAL(pos);
  if (topolow_fit_interruptible(&pb, &pr, &rs, poll_interrupt, NULL) != TOPOLOW_OK) { UNPROTECT(1); Rf_error("%s", rs.message); }
  const char* names[] = {"positions", "converged", "iterations", "final_mae", "final_k", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));                 /* src/optimization.cpp:375-381 */
  SET_VECTOR_ELT(out, 0, pos);
  SET_VECTOR_ELT(out, 1, Rf_ScalarLogical(rs.converged));
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(rs.iterations));
  SET_VECTOR_ELT(out, 3, Rf_ScalarReal(rs.final_mae));
  SET_VECTOR_ELT(out, 4, Rf_ScalarReal(rs.final_k));
  UNPROTECT(2);
  return out;
}
static const R_CallMethodDef CallEntries[] = {
  {"_topolow_optimize_layout_b200", (DL_FUNC)&_topolow_optimize_layout_b200, 16}, {NULL, NULL, 0}};
void R_init_topolowb200(DllInfo* dll) { R_registerRoutines(dll, NULL, CallEntries, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }
