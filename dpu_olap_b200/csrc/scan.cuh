// scan.cuh — internal (non-ABI) device-wide exclusive scan used by the partitioner.
#pragma once
#include "common.cuh"

size_t b2_scan_ws_bytes(int64_t n);
// d_out[i] = sum of d_in[0..i). d_in must be 16 B aligned. Enqueues on `s`, does not synchronise.
int b2_exclusive_scan_u32_u64(b2_ctx* ctx, const uint32_t* d_in, uint64_t* d_out, int64_t n,
                              void* d_ws, size_t ws_bytes, cudaStream_t s);
