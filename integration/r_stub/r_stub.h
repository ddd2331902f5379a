/* integration/r_stub/r_stub.h - what the harness sees of the stand-in runtime (test infrastructure). */
#ifndef TOPOLOW_R_STUB_H
#define TOPOLOW_R_STUB_H
#include <setjmp.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"
struct stub_state {
  jmp_buf toplevel;              /* where Rf_error / Rf_onintr / an unguarded interrupt land */
  char error_message[512];
  int errors, onintr_calls, raw_interrupt_jumps;
  int interrupt_checks, interrupt_after;   /* raise an interrupt at the interrupt_after-th check (0 = never) */
  int protect_depth, type_errors, allocations;
  int rng_open, rng_violations, rng_draws;
  const char* option_mode;
  R_CallMethodDef registered[8];
  int n_registered, dynamic_symbols;
};
extern struct stub_state stub;
SEXP stub_real(R_xlen_t n, const double* v);
SEXP stub_int(R_xlen_t n, const int* v);
SEXP stub_matrix(int nrow, int ncol, const double* v);
SEXP stub_list(R_xlen_t n);
const char* stub_name(SEXP x, R_xlen_t i);
#endif
