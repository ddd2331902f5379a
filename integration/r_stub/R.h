/* integration/r_stub/R.h - stand-in (see Rinternals.h): the RNG bracket of R_ext/Random.h. */
#ifndef TOPOLOW_R_STUB_R_H
#define TOPOLOW_R_STUB_R_H
#ifdef __cplusplus
extern "C" {
#endif
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
#ifdef __cplusplus
}
#endif
#endif
