// topolow_b200/csrc/tilepass_f64_p2.cu - ExactF64 instantiations (D = 1..16) of the production kernel,
// 2 point(s) per lane (64-point tiles).
#define TL_KP 2
#define POLICY ExactF64
#define REAL double
#define SUFFIX f64_
#include "tilepass_inst.inc"
