// partition.cuh — internal (non-ABI) interface of the radix partitioner (partition.cu), shared
// with the join (join.cu).
#pragma once
#include "common.cuh"

// Which bits of wang_hash(key) one partitioning pass looks at. Bit fields are taken from the
// most significant end, as the reference's radix bucket does (BUCKET_OF = hash >> shift,
// dpu/shared/kernels/partition.c:45-46):
//   [ skip | sel_bits | ... shl ... | bits | rest ]
//   selected(key)  <=>  sel_bits == 0  ||  ((h << sel_shl) >> (32 - sel_bits)) == sel_val
//   bucket(key)     =   (h << shl) >> (32 - bits)
struct PartGeom {
  int bits = 0;      // log2 fan-out of this pass, 0..10
  int shl = 0;       // hash bits consumed before this pass
  int sel_bits = 0;  // selection predicate (hash-space slice), 0 = take every row
  int sel_shl = 0;
  uint32_t sel_val = 0;
  // value predicate pushed down from a fused pipeline (filter -> join): keep a row iff its VALUE is
  // below val_thr. Rows that fail are dropped by this pass (counted by neither kernel).
  bool val_pred = false;
  uint32_t val_thr = 0;
#ifdef B2_LAB
  int lab = 0;  // tools/part_lab.py experiments (env B2_LAB_SCATTER): 1 = synthetic rows instead of loads, 2 = no flush
#endif
};

constexpr int kPartMaxBits = 10;           // per pass (warp-private u16 counters for 1024 bins)
constexpr int kPartTile = 8192;            // rows per tile (512 threads x 16 rows)
constexpr int kPartThreads = 512;

__host__ __device__ __forceinline__ uint32_t part_bucket(uint32_t h, int shl, int bits) {
  return bits == 0 ? 0u : ((h << shl) >> (32 - bits));
}
__host__ __device__ __forceinline__ bool part_selected(uint32_t h, int sel_shl, int sel_bits,
                                                       uint32_t sel_val) {
  return sel_bits == 0 || ((h << sel_shl) >> (32 - sel_bits)) == sel_val;
}

struct PartInput {
  // SoA: keys/vals separate columns (vals == nullptr: value = row index, as the reference's
  // selection_indices_vector, partition.c:267-294). AoS: pairs of (key, value).
  const uint32_t* keys = nullptr;
  const uint32_t* vals = nullptr;
  const uint2* pairs = nullptr;
};

// Workspace needed by one pass over n rows split into nseg segments with fan-out 2^bits.
size_t part_pass_ws_bytes(int64_t n, int64_t nseg, int bits);

// One partitioning pass. Rows of segment s are in[d_seg_off[s] .. d_seg_off[s+1]) and are
// partitioned independently into 2^bits buckets each; output positions are dense in
// (segment, bucket) order starting at row 0 of d_out. d_part_off receives nseg*2^bits + 1
// boundaries. out_cap: capacity of d_out in rows; rows beyond it are dropped and *d_overflow
// (nullable) is set to 1. Enqueues on s; no host synchronisation.
int part_pass(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
              const PartGeom& g, uint2* d_out, int64_t out_cap, int64_t* d_part_off,
              unsigned int* d_overflow, void* d_ws, size_t ws_bytes, cudaStream_t s);

// The two phases of part_pass, separately: count (unit table, histogram, scan, boundaries) and
// scatter. Between them the caller may turn the boundaries into per-bucket destination ADDRESSES
// (d_bucket_addr, 2^bits byte addresses, single segment only): the fused multi-GPU shuffle points
// them into the peers' receive buffers. Same n / segments / geometry / workspace for both calls.
// clamp_rows: boundaries written to d_part_off never exceed it (the capacity of the scatter's output).
int part_count(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
               const PartGeom& g, int64_t* d_part_off, void* d_ws, size_t ws_bytes, cudaStream_t s,
               int64_t clamp_rows = INT64_MAX);
int part_scatter(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
                 const PartGeom& g, uint2* d_out, int64_t out_cap, const uint64_t* d_bucket_addr,
                 unsigned int* d_overflow, void* d_ws, size_t ws_bytes, cudaStream_t s,
                 const int64_t* d_abort = nullptr);  // peer mode: *d_abort != 0 on the device = store nothing

// Full partitioning by `bits` hash bits (after discarding `shl`), in one pass (bits <= 10) or two
// (coarse pass + segmented fine pass). Result: d_out holds the rows grouped into 2^bits
// partitions, d_off (2^bits + 1 int64) their boundaries. d_tmp (capacity cap rows) is only used
// by the two-pass path. sel_*: hash-space slice predicate (see PartGeom).
size_t part_full_ws_bytes(int64_t n, int bits);
// val_pred / val_thr: value predicate applied by the first pass (see PartGeom).
int part_full(b2_ctx* ctx, const PartInput& in, int64_t n, int bits, int shl, int sel_shl,
              int sel_bits, uint32_t sel_val, uint2* d_out, uint2* d_tmp, int64_t cap,
              int64_t* d_off, unsigned int* d_overflow, void* d_ws, size_t ws_bytes, cudaStream_t s,
              bool val_pred = false, uint32_t val_thr = 0);
