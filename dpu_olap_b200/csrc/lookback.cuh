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

// ---- hierarchical counted sums ----------------------------------------------------------------
// The filter does not walk descriptors at all (filter.cu): every tile adds (1 << 48 | count) to
// the word of its group of 32 tiles and of its super-group of 1024 tiles with one fire-and-forget
// atomic each. A word whose contributor count (bits 63:48) has reached the group size holds the
// final sum (bits 47:0). The exclusive prefix of tile t is then
//     sum of full super-groups before t  (kept as a running base by the owning CTA)
//   + sum of the <= 31 full groups between the super-group boundary and t's group
//   + sum of the <= 31 tile counts of t's own group before t,
// i.e. two loads per lane, and — unlike a look-back chain — nothing a tile waits for depends on
// another tile's look-back having finished, only on earlier tiles having been COUNTED.
constexpr int kGroupShift = 5;         // 32 tiles per group
constexpr int kSuperShift = 10;        // 1024 tiles per super-group
constexpr uint64_t kSumMask = (1ull << 48) - 1;
constexpr uint32_t kAggFlag = 0x80000000u;  // per-tile count word: published flag | count

__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_relaxed_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Fire-and-forget relaxed gpu-scope add (no return value: compiles to RED, nothing to wait for).
__device__ __forceinline__ void red_add_relaxed_gpu_u64(uint64_t* p, uint64_t v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
