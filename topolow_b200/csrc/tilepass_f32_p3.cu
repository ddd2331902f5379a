// topolow_b200/csrc/tilepass_f32_p3.cu - FastF32 instantiations (D = 1..16) of the production kernel,
// 3 point(s) per lane (96-point tiles).
#define TL_KP 3
#define POLICY FastF32
#define REAL float
#define SUFFIX f32_
#include "tilepass_inst.inc"
