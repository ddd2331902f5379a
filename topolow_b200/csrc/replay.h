// topolow_b200/csrc/replay.h
#pragma once
#include "../../include/topolow_b200.h"
#include "common.cuh"

namespace tl {
int replay_max_n();
// Throws tl::CudaError; fills res (positions must be caller-allocated).
void run_replay(const topolow_problem& pb, const topolow_params& pr, topolow_result& res,
                topolow_interrupt_fn poll, void* user);
}  // namespace tl
