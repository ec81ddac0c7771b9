// filter64.cu — order-preserving `v < threshold` over 64-bit columns (uint64 / int64 / float64).
//
// SURVEY.md section 8f-3 ("other fixed-width types": the reference fixes `#define T uint32_t`,
// dpu/shared/common.h:3, and its kernel_filter, dpu/shared/kernels/filter.c:57-177, moves 4-byte
// items). Semantics are those of the reference's oracle, the Acero filter plan with
// less(field, literal) (host/filter/filter_native.cc:52-66) over a 64-bit column: rows keep their
// order, a null row is dropped, a NaN row is never selected, the result has no nulls.
//
// Two kernels compute the same result (ctx tunable B2_TUNE_FILTER64_KERNEL):
//   0  filter64_single_pass_kernel (default): persistent CTAs over 2048-row tiles (16 KB) dealt
//      round-robin. A CTA loads its tile (row pairs as 128-bit words), ranks the selected rows of each
//      warp with one packed shuffle scan, publishes the tile's count as counted sums (lookback.cuh:
//      groups of 32 and 1024 tiles, the scheme of the 32-bit kernel) and parks the selected rows
//      compacted in a shared-memory ring; a parked tile is written out as one contiguous run once its
//      global offset is computable, i.e. once every earlier tile has been COUNTED (the reference's
//      serial handshake between tasklets, filter.c:28-55, without the serialisation). Every row is
//      read once and every selected row written once: 8 + 8 s bytes per row (s = selectivity).
//      History and measurements: profiles/r2_filter64.md.
//   1  counted two-pass compaction: filter64_count_kernel (selected rows per tile), an exclusive scan
//      of the tile counts (csrc/scan.cu), filter64_compact_kernel (re-reads the tile and writes the
//      selected rows behind the tile's offset): 8 + 8 + 8 s bytes per row. Kept as the cross-check.
// Both leave the exclusive prefix of every tile (and the total as entry ntiles) in the workspace;
// filter64_end_kernel turns them into batch_end[b] = rows selected up to the end of batch b (one warp
// per batch recounts the part of the tile the boundary cuts) and the total.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "lookback.cuh"
#include "scan.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSlices = 8;                       // rows per thread and tile
constexpr int kTile = kThreads * kSlices;        // 2048 rows = 16 KB
constexpr int kWarps = kThreads / 32;
constexpr int kPrefetchTiles = 1;                // single-pass kernel: L2 request this many generations ahead

template <int kType>
__device__ __forceinline__ bool lt64(uint64_t a, uint64_t thr) {
  if (kType == B2_I64) return (long long)a < (long long)thr;
  if (kType == B2_F64) return __longlong_as_double((long long)a) < __longlong_as_double((long long)thr);
  return a < thr;
}

// A thread's rows of a tile: r0 + j * kStride, j = 0 .. 7. All loads are issued before the first value
// is looked at (eight independent 8-byte loads in flight per thread); rows past the end of the column
// are not loaded and never selected. `full`: the whole tile lies inside the column.
template <int kType, int kStride>
__device__ __forceinline__ void load_rows(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid,
                                          int64_t r0, int64_t n, bool full, uint64_t thr,
                                          uint64_t (&v)[kSlices], bool (&ok)[kSlices]) {
  uint32_t vb[kSlices];
  if (full) {
#pragma unroll
    for (int j = 0; j < kSlices; ++j) v[j] = ld_stream_u64(in + r0 + j * kStride);
    if (valid) {
#pragma unroll
      for (int j = 0; j < kSlices; ++j) vb[j] = valid[(r0 + j * kStride) >> 3];
    }
#pragma unroll
    for (int j = 0; j < kSlices; ++j) {
      ok[j] = lt64<kType>(v[j], thr);
      if (valid) ok[j] = ok[j] && ((vb[j] >> ((r0 + j * kStride) & 7)) & 1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kSlices; ++j) {
      const int64_t i = r0 + j * kStride;
      v[j] = i < n ? ld_stream_u64(in + i) : 0ull;
      vb[j] = (valid && i < n) ? valid[i >> 3] : 0xffu;
    }
#pragma unroll
    for (int j = 0; j < kSlices; ++j) {
      const int64_t i = r0 + j * kStride;
      ok[j] = i < n && lt64<kType>(v[j], thr) && ((vb[j] >> (i & 7)) & 1);
    }
  }
}
// The two-pass kernels' layout: thread t owns rows base + j * 256 + t.
template <int kType>
__device__ __forceinline__ void load_tile(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid,
                                          int64_t base, int64_t n, uint32_t tid, uint64_t thr,
                                          uint64_t (&v)[kSlices], bool (&ok)[kSlices]) {
  load_rows<kType, kThreads>(in, valid, base + tid, n, base + kTile <= n, thr, v, ok);
}

// Is row `i` of the packed column selected? (bit i of the packed validity bitmap, when there is one)
template <int kType>
__device__ __forceinline__ bool selected(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid,
                                         int64_t i, uint64_t thr, uint64_t* v) {
  *v = in[i];
  bool ok = lt64<kType>(*v, thr);
  if (valid) ok = ok && ((valid[i >> 3] >> (i & 7)) & 1);
  return ok;
}

template <int kType>
__global__ void __launch_bounds__(kThreads)
filter64_count_kernel(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid, int64_t n, uint64_t thr,
                      int64_t ntiles, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t wsum[kWarps];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * kTile;
    uint32_t c = 0;
    uint64_t v[kSlices];
    bool ok[kSlices];
    load_tile<kType>(in, valid, base, n, tid, thr, v, ok);
#pragma unroll
    for (int j = 0; j < kSlices; ++j) c += ok[j] ? 1u : 0u;
    c = warp_reduce_sum_u32(c);
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (tid == 0) {
      uint32_t s = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += wsum[w];
      tile_cnt[t] = s;
    }
    __syncthreads();
  }
}

template <int kType>
__global__ void __launch_bounds__(kThreads)
filter64_compact_kernel(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid, int64_t n, uint64_t thr,
                        int64_t ntiles, const uint64_t* __restrict__ tile_off, uint64_t* __restrict__ out) {
  __shared__ uint32_t cnt[kSlices * kWarps];  // selected rows of (slice j, warp w), in row order j * 8 + w
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = lanemask_lt();
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * kTile;
    uint64_t v[kSlices];
    bool ok[kSlices];
    uint32_t rank[kSlices];  // rank inside the warp's 32 rows of slice j, or ~0: not selected
    load_tile<kType>(in, valid, base, n, tid, thr, v, ok);
#pragma unroll
    for (int j = 0; j < kSlices; ++j) {
      const uint32_t m = __ballot_sync(0xffffffffu, ok[j]);
      rank[j] = ok[j] ? (uint32_t)__popc(m & lt) : 0xffffffffu;
      if (lane == 0) cnt[j * kWarps + warp] = (uint32_t)__popc(m);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 counts, two per lane
      const uint32_t a = cnt[2 * lane], b = cnt[2 * lane + 1];
      uint32_t incl = a + b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
      }
      const uint32_t excl = incl - a - b;
      cnt[2 * lane] = excl;
      cnt[2 * lane + 1] = excl + a;
    }
    __syncthreads();
    uint64_t* dst = out + tile_off[t];
#pragma unroll
    for (int j = 0; j < kSlices; ++j)
      if (rank[j] != 0xffffffffu) dst[cnt[j * kWarps + warp] + rank[j]] = v[j];
    __syncthreads();
  }
}

struct F64Head {
  unsigned long long ticket;  // CTA start ranks
  unsigned long long pad[7];
};

// Workspace words of the single-pass kernel (all zeroed per launch): the hierarchical counted sums of
// lookback.cuh, as the 32-bit kernel uses them (filter.cu).
struct F64Sums {
  uint32_t* agg;       // [ntiles]            kAggFlag | rows selected in the tile
  uint64_t* grp;       // [ntiles / 32 + 1]   contributors << 48 | rows selected in the group of 32 tiles
  uint64_t* sgrp;      // [ntiles / 1024 + 1] the same per super-group of 1024 tiles
  uint64_t* tile_off;  // [ntiles + 1]        out: rows selected before tile t; entry ntiles = the total
};

// Persistent CTAs, tiles dealt round-robin: the CTA that started v-th owns tiles v, v + G, v + 2G, ...
// (G = grid size <= the number of CTAs the device holds at once), so all CTAs work on the same
// generation of G consecutive tiles. Warp w of a CTA owns rows [256 w, 256 w + 256) of the tile in four
// segments of 64 rows; a lane owns two adjacent rows of each segment (one 128-bit load). Its per-segment
// counts (0..2) travel as the four bytes of ONE register through a single five-step shuffle scan, which
// ranks all 256 rows of the warp; the CTA-wide step is a sum over eight warp totals.
//
// A tile's global offset follows from the counted sums — the full super-groups before it (a running
// register), the <= 31 full groups before its group, the <= 31 tile counts before it in its group:
// two loads per lane — once every EARLIER tile has been counted. Nothing forces a CTA to wait for
// that: the selected rows of a counted tile sit compacted in a shared-memory ring (kRing rows) with a
// queue entry (ring position, rows), and a tile is RETIRED (offset computed, rows written out as one
// contiguous run) in a later iteration, whenever its offset has become computable. Only when the ring
// cannot take another full tile, or the queue is full, or the CTA has run out of tiles, does warp 0
// block for the oldest entry. At 25 % selected the ring holds eight tiles, so a CTA can run up to
// eight generations ahead of the slowest one; at 100 % it degrades to a lag of one tile.
//
// Iteration k of a CTA, two barriers:
//   A  issue the loads of tile k (registers); ask the copy engine for tile k + pf (L2 request)
//   B  warp 0: offsets of the queue entries that can (or must) be retired
//   C  write the retired tiles' rows out (under the latency of A's loads)
//   D  rank tile k's selected rows inside each warp (packed scan), post the warp totals
//   E  publish tile k's count (one store, two fire-and-forget adds), queue entry
//   F  stage tile k's selected rows in the ring
constexpr int kRing = 4096;  // rows of the staging ring (32 KB): at least one tile
constexpr int kQueue = 16;   // tiles a CTA may have counted but not written out
static_assert(kRing >= kTile && (kRing & (kRing - 1)) == 0 && (kQueue & (kQueue - 1)) == 0, "ring geometry");

struct F64Prefix {  // warp 0's running state
  uint64_t base_sg = 0;  // rows selected in the super-groups before `sg`
  int64_t sg = 0;
};
// Executed by warp 0: global offset of `tile` if every earlier tile has been counted (waits for that
// when `block`). Returns false when it has not (and !block).
__device__ __forceinline__ bool f64_tile_offset(const F64Sums& w, F64Prefix& st, int64_t tile, bool block,
                                                uint32_t lane, uint64_t* prefix) {
  while (st.sg < (tile >> kSuperShift)) {  // super-groups not folded into the base yet
    uint64_t sw = 0;
    if (lane == 0) {
      while (((sw = ld_relaxed_gpu_u64(w.sgrp + st.sg)) >> 48) != (1u << kSuperShift)) {
        if (!block) break;
        __nanosleep(64);
      }
    }
    sw = __shfl_sync(0xffffffffu, sw, 0);
    if ((sw >> 48) != (1u << kSuperShift)) return false;
    st.base_sg += sw & kSumMask;
    ++st.sg;
  }
  const int64_t g0 = st.sg << (kSuperShift - kGroupShift);  // first group of the super-group
  const int ng = (int)((tile >> kGroupShift) - g0);          // full groups before the tile's group
  const int64_t t0 = (tile >> kGroupShift) << kGroupShift;   // first tile of the tile's group
  const int na = (int)(tile - t0);                           // tiles of the group before the tile
  uint32_t ns = 32;
  while (true) {
    uint64_t gw = 0;
    uint32_t aw = kAggFlag;
    if ((int)lane < ng) gw = ld_relaxed_gpu_u64(w.grp + g0 + lane);
    if ((int)lane < na) aw = ld_relaxed_gpu_u32(w.agg + t0 + lane);
    const bool ready = ((int)lane >= ng || (gw >> 48) == (1u << kGroupShift)) && (aw & kAggFlag);
    if (__all_sync(0xffffffffu, ready)) {
      *prefix = st.base_sg + warp_reduce_sum_u64((gw & kSumMask) + (aw & ~kAggFlag));
      return true;
    }
    if (!block) return false;
    __nanosleep(ns);
    if (ns < 256) ns <<= 1;
  }
}

template <int kType>
__global__ void __launch_bounds__(kThreads)
filter64_single_pass_kernel(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid, int64_t n,
                            uint64_t thr, int64_t ntiles, F64Head* __restrict__ head, const F64Sums w,
                            uint64_t* __restrict__ out, int pf) {
  __shared__ uint64_t ring[kRing];       // selected rows of the counted tiles, compacted, in tile order
  __shared__ uint64_t q_prefix[kQueue];  // queue entry -> global offset (filled when retired)
  __shared__ uint32_t q_total[kQueue];   // queue entry -> rows
  __shared__ __align__(16) uint32_t wtot[kWarps];  // selected rows per warp of the tile being ranked
  __shared__ uint32_t s_nret;
  __shared__ uint32_t s_first;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_first = (uint32_t)atomicAdd(&head->ticket, 1ull);
  __syncthreads();
  const int64_t stride = gridDim.x;
  const int64_t first = s_first;
  const bool al16 = (reinterpret_cast<uintptr_t>(in) & 15) == 0;  // row pairs can be loaded as 128-bit words
  F64Prefix st;
  // uniform across the CTA: queue entries [qh, qt) are pending, entry e belongs to tile first + e * stride;
  // their rows occupy ring positions [r_tail, r_tail + used), oldest first
  uint32_t qh = 0, qt = 0, r_tail = 0, used = 0;
  // this thread's rows inside a tile: lrow + 64 j + e (j = 0..3 segments of 64 rows per warp, e = 0, 1)
  const uint32_t lrow = warp * (kTile / kWarps) + 2 * lane;
  for (int64_t k = 0;; ++k) {
    const int64_t tile = first + k * stride;
    const bool have = tile < ntiles;
    if (!have && qh == qt) break;
    uint64_t v[kSlices];     // v[2 j + e]
    uint32_t vbyte = 0xffu;  // lane l: byte l of the 32 validity bytes of the warp's 256 rows
    const int64_t row0 = tile * kTile;
    const uint32_t limit = (tile + 1) * kTile <= n ? (uint32_t)kTile : (uint32_t)(have ? n - row0 : 0);
    if (have) {  // A: loads only — nothing looks at the values before D
      if (warp == 0) {
        const int64_t next = tile + pf * stride;
        if (lane == 0 && pf > 0 && next < ntiles)
          l2_prefetch(in + next * kTile, (int64_t)sizeof(uint64_t) * min((int64_t)kTile, n - next * kTile));
      }
      const uint64_t* p = in + row0 + lrow;
      if (limit == kTile && al16) {
#pragma unroll
        for (int j = 0; j < kSlices / 2; ++j) {
          const uint4 q = ld_stream_v4(reinterpret_cast<const uint4*>(p + 64 * j));
          v[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
          v[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kSlices / 2; ++j) {
          v[2 * j] = lrow + 64 * j < limit ? ld_stream_u64(p + 64 * j) : 0ull;
          v[2 * j + 1] = lrow + 64 * j + 1 < limit ? ld_stream_u64(p + 64 * j + 1) : 0ull;
        }
      }
      if (valid) {
        const int64_t vb = ((row0 + warp * (kTile / kWarps)) >> 3) + lane;
        if (vb * 8 < n) vbyte = valid[vb];
      }
    }
    if (warp == 0) {  // B: retire what can be retired; block only for what must be
      __syncwarp();   // lane 0 pushed the newest queue entry
      uint32_t nret = 0, u = used;
      while (qh + nret != qt) {
        const uint32_t e = (qh + nret) & (kQueue - 1);
        const bool must = !have || u + kTile > kRing || qt - (qh + nret) == kQueue;
        const int64_t t = first + (int64_t)(qh + nret) * stride;
        uint64_t prefix;
        if (!f64_tile_offset(w, st, t, must, lane, &prefix)) break;
        const uint32_t total = q_total[e];
        if (lane == 0) {
          q_prefix[e] = prefix;
          w.tile_off[t] = prefix;
          if (t == ntiles - 1) w.tile_off[ntiles] = prefix + total;
        }
        u -= total;
        ++nret;
      }
      if (lane == 0) s_nret = nret;
    }
    __syncthreads();  // #1: retired offsets posted; the previous iteration's staging and queue entry are visible
    {                 // C
      const uint32_t nret = s_nret;
      for (uint32_t x = 0; x < nret; ++x) {
        const uint32_t e = (qh + x) & (kQueue - 1);
        const uint32_t total = q_total[e];
        uint64_t* __restrict__ dst = out + q_prefix[e] + tid;
#pragma unroll 1
        for (uint32_t i = tid; i < total; i += kThreads, dst += kThreads) *dst = ring[(r_tail + i) & (kRing - 1)];
        r_tail = (r_tail + total) & (kRing - 1);
        used -= total;
      }
      qh += nret;
    }
    // D: which of this lane's 8 rows are selected (bit 2 j + e), how many per segment (byte j of c: 0..2),
    // and ONE packed shuffle scan over the lanes for all four segments
    uint32_t okm = 0, pos = 0;
    if (have) {
#pragma unroll
      for (int x = 0; x < kSlices; ++x) okm |= lt64<kType>(v[x], thr) ? 1u << x : 0u;
      if (limit != kTile) {  // the column's last tile: rows past its end are not selected
#pragma unroll
        for (int x = 0; x < kSlices; ++x)
          if (lrow + 64 * (x >> 1) + (x & 1) >= limit) okm &= ~(1u << x);
      }
      if (valid) {  // rows 64 j + 2 lane + e of the warp's 256: bits 2 (lane & 3) + e of byte 8 j + lane / 4
        uint32_t vm = 0;
#pragma unroll
        for (int j = 0; j < kSlices / 2; ++j) {
          const uint32_t byte = __shfl_sync(0xffffffffu, vbyte, 8 * j + (lane >> 2));
          vm |= ((byte >> (2 * (lane & 3))) & 3u) << (2 * j);
        }
        okm &= vm;
      }
      const uint32_t pair = (okm & 0x55u) + ((okm >> 1) & 0x55u);  // 2-bit counts of the four segments
      const uint32_t c = (pair & 3u) | ((pair & 0xcu) << 6) | ((pair & 0x30u) << 12) | ((pair & 0xc0u) << 18);
      uint32_t incl = c;  // bytewise: no segment of 64 rows overflows a byte
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t seg = __shfl_sync(0xffffffffu, incl, 31);  // rows selected per segment (<= 64 each)
      // position of this lane's first selected row of segment j among the warp's selected rows:
      // rows of the segments before (<= 192) + rows of the lanes before in this segment (<= 62)
      pos = (incl - c) + ((seg << 8) + (seg << 16) + (seg << 24));
      if (lane == 0) wtot[warp] = (seg & 0xff) + ((seg >> 8) & 0xff) + ((seg >> 16) & 0xff) + (seg >> 24);
    }
    __syncthreads();  // #2: warp totals posted; the retired rows have left the ring
    if (have) {       // E, F
      uint32_t before = 0, total = 0;
      {
        static_assert(kWarps == 8, "two 128-bit loads");
        const uint4 a = *reinterpret_cast<const uint4*>(wtot), c = *reinterpret_cast<const uint4*>(wtot + 4);
        const uint32_t t[kWarps] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int x = 0; x < kWarps; ++x) {
          before += x < (int)warp ? t[x] : 0u;
          total += t[x];
        }
      }
      if (tid == 0) {
        q_total[qt & (kQueue - 1)] = total;
        st_relaxed_gpu_u32(w.agg + tile, kAggFlag | total);
        red_add_relaxed_gpu_u64(w.grp + (tile >> kGroupShift), (1ull << 48) | total);
        red_add_relaxed_gpu_u64(w.sgrp + (tile >> kSuperShift), (1ull << 48) | total);
      }
      const uint32_t at = r_tail + used + before;  // the ring's head + the share of the warps before this one
#pragma unroll
      for (int j = 0; j < kSlices / 2; ++j) {
        uint32_t at_j = at + ((pos >> (8 * j)) & 0xffu);
        if (okm & (1u << (2 * j))) ring[at_j++ & (kRing - 1)] = v[2 * j];
        if (okm & (2u << (2 * j))) ring[at_j & (kRing - 1)] = v[2 * j + 1];
      }
      used += total;
      ++qt;
    }
  }
}

// One warp per batch: rows selected in [0, batch_off[b + 1]) = offset of the tile the boundary falls
// into + the selected rows of that tile before the boundary. The last warp also writes the total.
template <int kType>
__global__ void __launch_bounds__(kThreads)
filter64_end_kernel(const uint64_t* __restrict__ in, const uint8_t* __restrict__ valid, int64_t n, uint64_t thr,
                    int64_t ntiles, const uint64_t* __restrict__ tile_off, const int64_t* __restrict__ batch_off,
                    int64_t nbatches, int64_t batch_len, int64_t* __restrict__ batch_end,
                    int64_t* __restrict__ total) {
  const int64_t b = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (b == 0 && lane == 0 && total) *total = (int64_t)tile_off[ntiles];
  if (b >= nbatches) return;
  const int64_t r = batch_off ? batch_off[b + 1] : (b + 1) * batch_len;  // first row after batch b
  const int64_t t = r / kTile;
  uint32_t c = 0;
  for (int64_t i = t * kTile + lane; i < r; i += 32) {
    uint64_t v;
    if (selected<kType>(in, valid, i, thr, &v)) ++c;
  }
  c = warp_reduce_sum_u32(c);
  if (lane == 0) batch_end[b] = (int64_t)tile_off[min(t, ntiles)] + c;
}

struct F64Layout {
  int64_t ntiles;
  // two-pass: tile counts | tile offsets | scan workspace; single pass: head | counted sums | tile offsets
  size_t off_cnt, off_scan, off_scanws, two_pass_total;
  size_t off_head, off_agg, off_grp, off_sgrp, off_tileoff, single_total;
  size_t total;
};
F64Layout f64_layout(int64_t n) {
  F64Layout L;
  L.ntiles = (n + kTile - 1) / kTile;
  size_t o = 0;
  L.off_cnt = o;    o += b2_align_up((size_t)(L.ntiles + 1) * 4, 256);
  L.off_scan = o;   o += b2_align_up((size_t)(L.ntiles + 1) * 8, 256);
  L.off_scanws = o; o += b2_align_up(b2_scan_ws_bytes(L.ntiles + 1), 256);
  L.two_pass_total = o;
  o = 0;
  L.off_head = o;    o += 256;
  L.off_agg = o;     o += b2_align_up((size_t)(L.ntiles + 1) * 4, 256);
  L.off_grp = o;     o += b2_align_up((size_t)((L.ntiles >> kGroupShift) + 1) * 8, 256);
  L.off_sgrp = o;    o += b2_align_up((size_t)((L.ntiles >> kSuperShift) + 1) * 8, 256);
  L.off_tileoff = o; o += b2_align_up((size_t)(L.ntiles + 1) * 8, 256);
  L.single_total = o;
  L.total = std::max(L.two_pass_total, L.single_total);
  return L;
}

template <int kType>
int filter64_impl(b2_ctx* ctx, const uint64_t* d_in, const uint8_t* d_valid, int64_t n, uint64_t thr,
                  const int64_t* d_batch_off, int64_t nbatches, int64_t batch_len, uint64_t* d_out,
                  int64_t* d_batch_end, int64_t* d_total, void* d_ws, size_t ws_bytes, cudaStream_t s) {
  const F64Layout L = f64_layout(n);
  if (ws_bytes < L.total) return b2_set_error(ctx, B2_ERR_WORKSPACE, "64-bit filter", "use b2_filter_64_ws_bytes()");
  B2_REQUIRE(ctx, L.ntiles < (1ll << 31), "column too large for one launch");
  char* base = static_cast<char*>(d_ws);
  const uint64_t* off = nullptr;  // exclusive prefix per tile, entry ntiles = the total
  if (ctx->tune[B2_TUNE_FILTER64_KERNEL] == 0) {
    F64Head* head = reinterpret_cast<F64Head*>(base + L.off_head);
    F64Sums w;
    w.agg = reinterpret_cast<uint32_t*>(base + L.off_agg);
    w.grp = reinterpret_cast<uint64_t*>(base + L.off_grp);
    w.sgrp = reinterpret_cast<uint64_t*>(base + L.off_sgrp);
    w.tile_off = reinterpret_cast<uint64_t*>(base + L.off_tileoff);
    uint64_t* tile_off = w.tile_off;
    B2_CUDA_OK(ctx, cudaMemsetAsync(base, 0, L.single_total, s));  // start ranks, counted sums, total of an empty column
    if (L.ntiles > 0) {
      // every CTA waits for tiles of the others: the grid must be resident as a whole
      static const int seen = b2_new_site();
      if (b2_first_use_on_device(ctx, seen)) {
        int per_sm = 0;
        B2_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter64_single_pass_kernel<kType>,
                                                                      kThreads, 0));
        if (per_sm < 1) return b2_set_error(ctx, B2_ERR_CUDA, "64-bit filter kernel", "does not fit an SM");
        ctx->site_value[(size_t)seen] = per_sm * ctx->sm_count;
      }
      int64_t grid = std::min<int64_t>(L.ntiles, ctx->site_value[(size_t)seen]);
      int pf = kPrefetchTiles;
#ifdef B2_LAB
      if (const char* e = getenv("B2_LAB_F64")) {  // "prefetch distance,CTAs per SM" (tools/filter64_lab.py)
        int ctas = 0;
        if (sscanf(e, "%d,%d", &pf, &ctas) == 2 && ctas > 0) grid = std::min<int64_t>(grid, (int64_t)ctas * ctx->sm_count);
      }
#endif
      filter64_single_pass_kernel<kType><<<(unsigned)grid, kThreads, 0, s>>>(d_in, d_valid, n, thr, L.ntiles, head, w,
                                                                             d_out, pf);
      B2_LAUNCH_CHECK(ctx, "filter64_single_pass_kernel");
    }
    off = tile_off;
  } else {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(base + L.off_cnt);
    uint64_t* scanned = reinterpret_cast<uint64_t*>(base + L.off_scan);
    B2_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, (size_t)(L.ntiles + 1) * 4, s));
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(L.ntiles, (int64_t)ctx->sm_count * 8));
    if (L.ntiles > 0) {
      filter64_count_kernel<kType><<<grid, kThreads, 0, s>>>(d_in, d_valid, n, thr, L.ntiles, cnt);
      B2_LAUNCH_CHECK(ctx, "filter64_count_kernel");
    }
    // ntiles + 1 entries: the last one is the total
    B2_RETURN_NOT_OK(b2_exclusive_scan_u32_u64(ctx, cnt, scanned, L.ntiles + 1, base + L.off_scanws,
                                               ws_bytes - L.off_scanws, s));
    if (L.ntiles > 0) {
      filter64_compact_kernel<kType><<<grid, kThreads, 0, s>>>(d_in, d_valid, n, thr, L.ntiles, scanned, d_out);
      B2_LAUNCH_CHECK(ctx, "filter64_compact_kernel");
    }
    off = scanned;
  }
  const int64_t warps = std::max<int64_t>(nbatches, 1);
  filter64_end_kernel<kType><<<(unsigned)((warps + kWarps - 1) / kWarps), kThreads, 0, s>>>(
      d_in, d_valid, n, thr, L.ntiles, off, d_batch_off, nbatches, batch_len, d_batch_end, d_total);
  B2_LAUNCH_CHECK(ctx, "filter64_end_kernel");
  return B2_OK;
}

}  // namespace

extern "C" {

size_t b2_filter_64_ws_bytes(int64_t n) { return n < 0 ? 0 : f64_layout(n).total; }

int b2_filter_lt_64_dev(b2_ctx* ctx, const void* d_in, int dtype, uint64_t threshold_bits, const uint8_t* d_valid,
                        int64_t n, const int64_t* d_batch_off, int64_t nbatches, int64_t batch_len, void* d_out,
                        int64_t* d_batch_end, int64_t* d_total, void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0 && nbatches >= 0 && batch_len >= 0, "negative size");
  B2_REQUIRE(ctx, dtype == B2_U64 || dtype == B2_I64 || dtype == B2_F64, "dtype must be B2_U64, B2_I64 or B2_F64");
  B2_REQUIRE(ctx, n == 0 || (d_in && d_out), "null column");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 7) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 7) == 0,
             "64-bit columns must be 8-byte aligned");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end, "d_batch_end is null");
  B2_REQUIRE(ctx, d_batch_off != nullptr || nbatches * batch_len == n, "uniform batches must cover the column");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "workspace must be 256 B aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t* in = static_cast<const uint64_t*>(d_in);
  uint64_t* out = static_cast<uint64_t*>(d_out);
  switch (dtype) {
    case B2_I64:
      return filter64_impl<B2_I64>(ctx, in, d_valid, n, threshold_bits, d_batch_off, nbatches, batch_len, out,
                                   d_batch_end, d_total, d_ws, ws_bytes, s);
    case B2_F64:
      return filter64_impl<B2_F64>(ctx, in, d_valid, n, threshold_bits, d_batch_off, nbatches, batch_len, out,
                                   d_batch_end, d_total, d_ws, ws_bytes, s);
    default:
      return filter64_impl<B2_U64>(ctx, in, d_valid, n, threshold_bits, d_batch_off, nbatches, batch_len, out,
                                   d_batch_end, d_total, d_ws, ws_bytes, s);
  }
}

}  // extern "C"
