// lookback.cuh — decoupled look-back over single-word tile descriptors (shared by the filter's
// stream compaction and the partitioner's histogram scan).
//
// A descriptor is one 64-bit word: bits 63:62 = status (0 not published, 1 = tile aggregate,
// 2 = inclusive running total), bits 61:0 = value. A relaxed gpu-scope 64-bit access is atomic,
// so status and value are always seen together and no fence is needed. Tile ids must be handed
// out in launch order (atomic ticket) so that every predecessor of a waiting tile is resident.
#pragma once
#include "common.cuh"

constexpr uint64_t kFlagAgg = 1ull << 62;  // descriptor holds this tile's own count
constexpr uint64_t kFlagInc = 2ull << 62;  // descriptor holds the inclusive running count
constexpr uint64_t kValMask = (1ull << 62) - 1;

// Executed by one full warp. Publishes this tile's aggregate, walks back over the predecessors
// 32 at a time until an inclusive total is found, publishes the tile's own inclusive total and
// returns the exclusive prefix of `tile` (in all lanes). carry_in (nullable) seeds tile 0.
__device__ __forceinline__ uint64_t lookback(uint64_t* __restrict__ desc, int64_t tile,
                                             uint64_t own_count, const int64_t* __restrict__ carry_in) {
  const uint32_t lane = lane_id();
  if (tile == 0) {
    const uint64_t carry = carry_in ? (uint64_t)*carry_in : 0ull;  // rows emitted by earlier calls
    if (lane == 0) st_relaxed_gpu_u64(desc, kFlagInc | (carry + own_count));
    return carry;
  }
  if (lane == 0) st_relaxed_gpu_u64(desc + tile, kFlagAgg | own_count);
  uint64_t excl = 0;
  int64_t pos = tile - 1;  // lane 0 inspects the nearest predecessor
  while (true) {
    const int64_t idx = pos - lane;
    uint64_t d = kFlagInc;  // virtual tile -1: inclusive count 0
    if (idx >= 0) d = ld_relaxed_gpu_u64(desc + idx);
    const uint32_t status = (uint32_t)(d >> 62);
    const uint32_t invalid = __ballot_sync(0xffffffffu, status == 0);
    const uint32_t inc = __ballot_sync(0xffffffffu, status == 2);
    uint32_t need = 0xffffffffu;  // lanes whose value we must add
    if (inc) need = (2u << (__ffs(inc) - 1)) - 1u;
    if (invalid & need) {
      __nanosleep(20);
      continue;  // some predecessor in the window has not published yet
    }
    uint64_t v = ((need >> lane) & 1u) ? (d & kValMask) : 0ull;
    excl += warp_reduce_sum_u64(v);
    if (inc) break;
    pos -= 32;
  }
  if (lane == 0) st_relaxed_gpu_u64(desc + tile, kFlagInc | (excl + own_count));
  return excl;
}
