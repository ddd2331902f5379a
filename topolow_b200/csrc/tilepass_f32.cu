// topolow_b200/csrc/tilepass_f32.cu - FP32 instantiations of the production kernel (D = 1..16).
#include "tilepass_launch.h"
#define POLICY FastF32
#define REAL float
#define SUFFIX f32
#include "tilepass_inst.inc"
