// topolow_b200/csrc/tilepass_f64_p1.cu - ExactF64 instantiations (D = 1..16) of the production kernel,
// 1 point(s) per lane (32-point tiles).
#define TL_KP 1
#define POLICY ExactF64
#define REAL double
#define SUFFIX f64_
#include "tilepass_inst.inc"
