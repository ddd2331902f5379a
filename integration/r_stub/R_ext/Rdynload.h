/* integration/r_stub/R_ext/Rdynload.h - stand-in (see ../Rinternals.h): routine registration. */
#ifndef TOPOLOW_R_STUB_RDYNLOAD_H
#define TOPOLOW_R_STUB_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct DllInfo_ DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* fortran, const void* external);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif
