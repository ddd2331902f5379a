// topolow_b200/csrc/rowblock.h - host interface of the row-block ("owner computes") mode, rowblock.cu.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/topolow_b200.h"
#include "common.cuh"

namespace tl {

constexpr int kMaxShards = 16;          // replicas of one map (GPUs of one NVSwitch domain)
constexpr int kRowTile = 256;           // rows of one repulsion CTA; slot count and shard bounds are multiples of it
constexpr int kChunk = 2048;            // partner slots of one repulsion work item (fixed: the summation tree of a
                                        // point's repulsion must not depend on the number of shards)
constexpr int kRedRows = 64;            // rows behind one reduction entry (one spring / MAE CTA)

struct RowPlan;

RowPlan* row_create(const topolow_problem& pb, const topolow_params& pr);   // throws BadArg / CudaError
void row_destroy(RowPlan* rp);
// Up to n_iters further iterations (stops on convergence); returns CUDA-event milliseconds.
double row_run(RowPlan& rp, int n_iters, cudaStream_t stream, topolow_interrupt_fn poll, void* user, bool* interrupted);
void row_result(RowPlan& rp, topolow_result& res, bool interrupted);
// {slots, ndim, stride, shards, rank, row0, own_rows, chunks, records, launches, iterations_launched, exchange_bytes_per_iter}
void row_info(const RowPlan& rp, int64_t* out, int cap);

// ---- several replicas of one map (one per GPU): see include/topolow_b200.h, topolow_shard_* ----
size_t row_ipc_blob_size();
void row_ipc_export(RowPlan& rp, void* blob);                          // this replica's shared block
void row_ipc_attach(RowPlan& rp, const void* blobs, int n_blobs);      // blobs of all ranks, rank order
void row_attach_local(RowPlan* const* plans, int n);                   // all replicas in this process (same device)
double row_run_lockstep(RowPlan* const* plans, int n, int n_iters);    // emulation: one stream, ranks in turn

}  // namespace tl
