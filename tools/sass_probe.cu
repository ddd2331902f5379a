// tools/sass_probe.cu - one instantiation of the production kernel for quick SASS studies:
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -c tools/sass_probe.cu -o /tmp/probe.o -I topolow_b200/csrc
// python tools/sass_sched.py /tmp/probe.o Li16 [--dump]
#define TL_KP 3
#include "tilepass.cuh"
namespace tl { namespace p3 {
template __global__ void tile_kernel<16, FastF32>(TileDev<float>, Geometry, FitParams, int, volatile int*);
}}
