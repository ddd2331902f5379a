// topolow_b200/csrc/tilepass_f64.cu - exact-FP64 instantiations of the production kernel (D = 1..16).
#include "tilepass_launch.h"
#define POLICY ExactF64
#define REAL double
#define SUFFIX f64
#include "tilepass_inst.inc"
