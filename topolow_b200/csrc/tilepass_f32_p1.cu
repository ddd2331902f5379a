// topolow_b200/csrc/tilepass_f32_p1.cu - FastF32 instantiations (D = 1..16) of the production kernel,
// 1 point(s) per lane (32-point tiles).
#define TL_KP 1
#define POLICY FastF32
#define REAL float
#define SUFFIX f32_
#include "tilepass_inst.inc"
