// topolow_b200/csrc/tilepass_f64_p3.cu - ExactF64 instantiations (D = 1..16) of the production kernel,
// 3 point(s) per lane (96-point tiles).
#define TL_KP 3
#define POLICY ExactF64
#define REAL double
#define SUFFIX f64_
#include "tilepass_inst.inc"
