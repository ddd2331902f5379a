// topolow_b200/csrc/edges.h
#pragma once
#include <vector>

#include "../../include/topolow_b200.h"
#include "tiledev.h"

namespace tl {
// Uploads the COO edge list of `pb`, returns device arrays: records sorted by tile-pair bucket
// (deterministic order inside a bucket) and the T*T+1 bucket offsets.  Throws CudaError /
// std::invalid_argument("edge index out of range").
void build_buckets(const topolow_problem& pb, const std::vector<int32_t>& slot_of_point, int T, int tile_points,
                   cudaStream_t stream,
                   EdgeRec** edges_out, uint32_t** bucket_off_out);
}  // namespace tl
