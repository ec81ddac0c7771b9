// filter.cu — order-preserving selection `v < threshold` over a uint32 column (stream compaction).
//
// Replaces the reference's DPU filter program (dpu/shared/kernels/filter.c:57-177): there each of
// 16 tasklets compacts a 128-element block in WRAM and a serial handshake chain
// (handshake_sync, filter.c:28-55) hands the running output offset from tasklet to tasklet, with
// a `write_carry` fix-up for 8-byte MRAM alignment (:109-131) and two barriers per round. The
// predicate is `item < (1 << 30)` (filter.c:25; Acero side: filter_native.cc:59).
//
// B200 design: ONE pass over the column (4 B read per row + 4 B written per selected row), run by
// a persistent, warp-specialised kernel (a few CTAs per SM, each looping over tiles):
//   * tiles are dealt round-robin: the CTA that started v-th (an atomic ticket taken once per CTA)
//     owns tiles v, v+G, v+2G, ... (G = grid size). All CTAs therefore work on the same
//     "generation" of G consecutive tiles at the same time, which keeps every look-back short.
//     (Dealing TILE tickets at prefetch time instead looked natural and ran 18x slower: a CTA then
//     holds several consecutive tiles it will only reach iterations later, and every successor's
//     look-back convoys behind it.)
//   * PRODUCER warp (one lane): streams the CTA's tiles into a ring of shared-memory stages with
//     TMA bulk copies (cp.async.bulk + mbarrier complete_tx). Bytes in flight per SM =
//     (stages - 1 - lag) x tile x CTAs, independent of what the compute warps are doing.
//   * COMPUTE warps: read their 16 rows per thread from shared memory (128-bit, conflict free),
//     count matches four segments at a time in ONE packed register (4 x 8-bit counters), so the
//     ranks of a warp's 512 rows cost a single 5-step shuffle scan; one CTA barrier per tile gives
//     the warp bases; selected rows are compacted IN PLACE into the stage buffer and written out
//     with fully coalesced stores `lag` iterations later, when the tile's global offset is known.
//   * Waiting: data arrival (TMA) is the only thing polled on an mbarrier. "Prefix posted" and
//     "stage drained" go through named hardware barriers (bar.arrive by the signalling warps,
//     bar.sync by the waiting ones), so a waiting warp is descheduled instead of spinning: the
//     mbarrier try_wait loops of the first version were 24 % of all executed instructions.
//   * PREFIX warp: turns tile counts into global output offsets WITHOUT a look-back chain. Each
//     tile adds its count to the word of its group (32 tiles) and super-group (1024 tiles) with a
//     fire-and-forget atomic; a word is final when its contributor count is complete. A tile's
//     offset = running sum of full super-groups (kept in a register by the owning CTA) + <= 31
//     full groups + <= 31 tile counts of its own group: two loads per lane, and nothing it waits
//     for depends on another tile's prefix being finished, only on earlier tiles having been
//     counted. The warp runs `lag` tiles behind the compute warps, which hides its L2 round trip.
//     (History, profiles/r1_filter.md: a classic decoupled look-back, 32 descriptors per round
//     from inside the compute path, capped the first version at 30 % of HBM peak; a 256-wide
//     look-back in a separate warp reached 45 % — the chain through "inclusive" descriptors kept
//     falling generations behind.)
// Tiles never straddle a batch boundary, so the inclusive count of the last tile of batch b is
// the end offset of result chunk b (FilterDpu::GetResult returns one chunk per input batch,
// host/filter/filter_dpu.cc:89-96,162-166) — a tiny second kernel collects those.
//
// Nullable columns (the DPU path has none, it passes a nullptr bitmap,
// filter_dpu.cc:91): Arrow's filter drops rows whose predicate is null, so a row is selected iff it
// is valid AND v < threshold. The kNullable variant reads the packed validity bitmap (bit i = row i,
// LSB first, as Arrow) — one 32-bit word per 8 lanes and segment — and turns null rows into
// 0xffffffff, the value that never satisfies `< threshold`; counting and compaction are unchanged.
#include "common.cuh"
#include "lookback.cuh"
#include "tma.cuh"

namespace {

constexpr int kVecPerThread = 4;  // uint4 per compute thread and tile (16 rows)

struct FilterWs {  // header of the caller workspace; zeroed with the count words on every launch
  unsigned long long ticket;       // CTA start ranks (static dealing) or tile tickets (dynamic)
  unsigned long long pad[7];
};

struct FilterArgs {
  const uint32_t* in;
  const uint32_t* valid;  // kNullable: validity bitmap over the packed rows, 4-byte words
  uint32_t* out;
  uint32_t thr;
  int64_t ntiles;
  int64_t batch_len, tiles_per_batch;  // uniform layout
  const int64_t* tile_row0;            // ragged layout: per-tile geometry (filter_geom_kernel)
  const int32_t* tile_len;
  const int64_t* carry_in;
  FilterWs* ws;
  uint32_t* agg;    // [ntiles]        kAggFlag | rows selected in the tile
  uint64_t* grp;    // [ntiles >> 5]   contributors << 48 | rows selected in the group of 32 tiles
  uint64_t* sgrp;   // [ntiles >> 10]  same for the super-group of 1024 tiles
  uint64_t* incl;   // [ntiles]        out: rows selected in tiles 0..t (+ carry), plain values
#ifdef B2_LAB
  int debug;  // lab build only (tools/filter_lab.py): 1 = no prefix (WRONG offsets), 2 = no TMA, 4 = dynamic tile tickets
#else
  static constexpr int debug = 0;  // the product build has no such switches: the branches fold away
#endif
};

// Compile-time shape of one kernel variant.
template <int kComputeThreads_, int kStages_, int kCtasPerSm_, int kLag_, bool kNamedBars_ = false>
struct FilterCfg {
  // kNamedBars: "prefix posted" and "stage drained" are signalled through named hardware barriers
  // (bar.arrive / bar.sync, ids 2 .. 2 + 2 * stages - 1) instead of mbarriers, so the waiting warps
  // are descheduled instead of polling (the polling loops were 24 % of all executed instructions)
  static constexpr bool kNamedBars = kNamedBars_;
  static_assert(!kNamedBars_ || 2 + 2 * kStages_ <= 16, "named barrier ids");
  static constexpr int kLag = kLag_;  // iterations between compacting a tile and writing it out
  static_assert(kLag_ >= 1 && kStages_ >= kLag_ + 2, "need a stage in flight besides the held ones");
  static constexpr int kComputeThreads = kComputeThreads_;
  static constexpr int kStages = kStages_;
  static constexpr int kCtasPerSm = kCtasPerSm_;
  static constexpr int kWarps = kComputeThreads / 32;
  static constexpr int kThreads = kComputeThreads + 64;  // + producer warp + look-back warp
  static constexpr int kTile = kComputeThreads * kVecPerThread * 4;  // rows
  static constexpr int kStageBytes = kTile * 4;
  // dynamic shared memory: stages | per-stage info | barriers | warp totals
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  // nullable variant: + one tile of validity bits per stage, filled by the same TMA producer
  static constexpr int kValidBytes = kTile / 8;
  static constexpr int kSmemBytesNullable = kSmemBytes + kStages * kValidBytes;
};

struct StageInfo {  // written by the producer before it completes full[s]
  int64_t tile;     // -1: no more tiles
  int64_t row0;
  int32_t len;
  int32_t tma;      // bit 0: the rows are in the stage buffer (else compute warps load them from
                    // global); bit 1 (nullable): so are the tile's validity bits
};

constexpr int kMaxStages = 8;
struct FilterSmemCtl {
  StageInfo info[kMaxStages];
  uint64_t full[kMaxStages];    // producer -> everyone: stage filled (TMA complete_tx or plain arrive)
  uint64_t empty[kMaxStages];   // compute warps -> producer: stage drained (count = kWarps)
  uint64_t tot[kMaxStages];     // compute thread 0 -> look-back warp: tile total posted
  uint64_t pre[kMaxStages];     // look-back warp -> compute warps: exclusive prefix posted
  uint32_t total[kMaxStages];
  uint64_t prefix[kMaxStages];
  uint32_t wtot[2][32];
  uint32_t vcta;                // this CTA's start rank
};
static_assert(sizeof(FilterSmemCtl) <= 1024, "control block must fit its reservation");

// The column's 32-bit type decides how `v < threshold` compares (the reference
// fixes `#define T uint32_t`, dpu/shared/common.h:3). Values are moved as raw 32-bit words.
enum { kCmpU32 = B2_U32, kCmpS32 = B2_I32, kCmpF32 = B2_F32 };
// A bit pattern that is never `< threshold` for any threshold: stands in for rows past the end of a
// tile and for null rows. (f32: a NaN — every ordered comparison with it is false.)
template <int kCmp>
__device__ __forceinline__ constexpr uint32_t never_lt() {
  return kCmp == kCmpU32 ? 0xffffffffu : (kCmp == kCmpS32 ? 0x7fffffffu : 0x7fc00000u);
}
// 1 if a < b in the column's type else 0.
template <int kCmp>
__device__ __forceinline__ uint32_t lt_32(uint32_t a, uint32_t b) {
  if (kCmp == kCmpS32) return (int32_t)a < (int32_t)b ? 1u : 0u;
  if (kCmp == kCmpF32) return __uint_as_float(a) < __uint_as_float(b) ? 1u : 0u;
  return a < b ? 1u : 0u;
}

// acc += inc if v < thr: compare + predicated add, two instructions per row.
template <int kCmp>
__device__ __forceinline__ void count_lt(uint32_t& acc, uint32_t v, uint32_t thr, uint32_t inc) {
  if (kCmp == kCmpS32) {
    asm("{\n\t.reg .pred q;\n\t"
        "setp.lt.s32 q, %1, %2;\n\t"
        "@q add.u32 %0, %0, %3;\n\t}"
        : "+r"(acc)
        : "r"(v), "r"(thr), "r"(inc));
  } else if (kCmp == kCmpF32) {
    asm("{\n\t.reg .pred q;\n\t.reg .f32 fa, fb;\n\t"
        "mov.b32 fa, %1;\n\tmov.b32 fb, %2;\n\t"
        "setp.lt.f32 q, fa, fb;\n\t"
        "@q add.u32 %0, %0, %3;\n\t}"
        : "+r"(acc)
        : "r"(v), "r"(thr), "r"(inc));
  } else {
    asm("{\n\t.reg .pred q;\n\t"
        "setp.lt.u32 q, %1, %2;\n\t"
        "@q add.u32 %0, %0, %3;\n\t}"
        : "+r"(acc)
        : "r"(v), "r"(thr), "r"(inc));
  }
}

template <typename Cfg, bool kNullable, int kCmp>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kCtasPerSm)
filter_lt_u32_kernel(const FilterArgs a) {
  constexpr int kCT = Cfg::kComputeThreads, kS = Cfg::kStages, kW = Cfg::kWarps, kTile = Cfg::kTile;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint32_t* const bufs = reinterpret_cast<uint32_t*>(smem_raw);
  FilterSmemCtl& ctl = *reinterpret_cast<FilterSmemCtl*>(smem_raw + kS * Cfg::kStageBytes);
  uint32_t* const vbufs = reinterpret_cast<uint32_t*>(smem_raw + Cfg::kSmemBytes);  // kNullable only

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      mbar_init(&ctl.full[s], 1);
      mbar_init(&ctl.empty[s], kW);
      mbar_init(&ctl.tot[s], 1);
      mbar_init(&ctl.pre[s], 1);
    }
    fence_mbar_init();
    ctl.vcta = (a.debug & 4) ? 0u : (uint32_t)atomicAdd(&a.ws->ticket, 1ull);
  }
  __syncthreads();
  const int64_t first_tile = ctl.vcta, tile_stride = gridDim.x;

  if (warp == kW) {
    // ================================ producer ================================
    // Lane 0 does the work; with named barriers the whole warp takes part in the waits.
    const uint64_t policy = l2_evict_first_policy();
    for (uint32_t n = 0;; ++n) {  // iterations of one CTA: far below 2^32
      const int s = (int)(n % kS);
      if (n >= (uint32_t)kS) {
        if (Cfg::kNamedBars) named_bar_sync(2 + kS + s, kCT + 32);
        else if (lane == 0) mbar_wait(&ctl.empty[s], ((n / kS) - 1) & 1);
      }
      int done = 0;
      if (lane == 0) {
        const int64_t tile = (a.debug & 4) ? (int64_t)atomicAdd(&a.ws->ticket, 1ull)
                                           : first_tile + (int64_t)n * tile_stride;
        StageInfo& si = ctl.info[s];
        if (tile >= a.ntiles) {
          si.tile = -1;
          mbar_arrive(&ctl.full[s]);
          done = 1;
        } else {
          int64_t row0;
          int32_t len;
          if (a.tile_row0) {
            row0 = a.tile_row0[tile];
            len = a.tile_len[tile];
          } else {
            const int64_t b = tile / a.tiles_per_batch, k = tile - b * a.tiles_per_batch;
            row0 = b * a.batch_len + k * kTile;
            const int64_t rest = a.batch_len - k * kTile;
            len = rest < kTile ? (int32_t)rest : kTile;
          }
          const uint32_t* src = a.in + row0;
          const bool tma = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((len & 3) == 0) &&
                           !(a.debug & 2);
          // the tile's 512 bytes of validity ride along when they start on a 16-byte boundary
          const bool vtma = kNullable && tma && len == kTile && (row0 & 127) == 0 &&
                            (reinterpret_cast<uintptr_t>(a.valid) & 15) == 0;
          si.tile = tile;
          si.row0 = row0;
          si.len = len;
          si.tma = (tma ? 1 : 0) | (vtma ? 2 : 0);
          if (tma) {
            fence_proxy_async();  // the stage was last touched through the generic proxy
            mbar_arrive_expect_tx(&ctl.full[s], (uint32_t)len * 4u + (vtma ? (uint32_t)Cfg::kValidBytes : 0u));
            tma_load_1d(bufs + (size_t)s * kTile, src, (uint32_t)len * 4u, &ctl.full[s], policy);
            if (vtma)
              tma_load_1d(vbufs + (size_t)s * (Cfg::kValidBytes / 4), a.valid + (row0 >> 5), Cfg::kValidBytes,
                          &ctl.full[s], policy);
          } else {
            mbar_arrive(&ctl.full[s]);
          }
        }
      }
      if (__shfl_sync(0xffffffffu, done, 0)) break;
    }
    return;
  }

  if (warp == kW + 1) {
    // ================================ prefix warp ================================
    uint64_t base = a.carry_in ? (uint64_t)*a.carry_in : 0ull;  // rows before super-group `sg`
    int64_t sg = 0;
    for (uint32_t n = 0;; ++n) {
      const int s = (int)(n % kS);
      const uint32_t par = (n / kS) & 1;
      mbar_wait(&ctl.full[s], par);
      const int64_t tile = ctl.info[s].tile;
      if (tile < 0) break;
      mbar_wait(&ctl.tot[s], par);
      const uint32_t total = ctl.total[s];
      uint64_t prefix;
      if (a.debug & 1) {
        prefix = (uint64_t)tile * kTile;
      } else {
        // full super-groups this CTA has not folded into its base yet (at most one per iteration
        // with static dealing, since the grid is smaller than a super-group)
        while (sg < (tile >> kSuperShift)) {
          uint64_t w = 0;
          if (lane == 0) {
            while (((w = ld_relaxed_gpu_u64(a.sgrp + sg)) >> 48) != (1u << kSuperShift)) __nanosleep(100);
          }
          base += __shfl_sync(0xffffffffu, w, 0) & kSumMask;
          ++sg;
        }
        const int64_t g0 = sg << (kSuperShift - kGroupShift);   // first group of the super-group
        const int ng = (int)((tile >> kGroupShift) - g0);        // full groups before tile's group
        const int64_t t0 = (tile >> kGroupShift) << kGroupShift; // first tile of tile's group
        const int na = (int)(tile - t0);                         // tiles of the group before tile
        uint32_t ns = 32;
        while (true) {
          uint64_t gw = 0;
          uint32_t aw = kAggFlag;
          if ((int)lane < ng) gw = ld_relaxed_gpu_u64(a.grp + g0 + lane);
          if ((int)lane < na) aw = ld_relaxed_gpu_u32(a.agg + t0 + lane);
          const bool ok = ((int)lane >= ng || (gw >> 48) == (1u << kGroupShift)) && (aw & kAggFlag);
          if (__all_sync(0xffffffffu, ok)) {
            prefix = base + warp_reduce_sum_u64((gw & kSumMask) + (aw & ~kAggFlag));
            break;
          }
          __nanosleep(ns);
          if (ns < 256) ns <<= 1;
        }
      }
      if (lane == 0) {
        a.incl[tile] = prefix + total;
        ctl.prefix[s] = prefix;
        if (!Cfg::kNamedBars) mbar_arrive(&ctl.pre[s]);
      }
      __syncwarp();
      if (Cfg::kNamedBars) named_bar_arrive(2 + s, kCT + 32);  // after lane 0's store (PTX producer pattern)
    }
    return;
  }

  // ================================ compute warps ================================
  for (uint32_t n = 0;; ++n) {  // 32-bit: the per-tile index arithmetic is on the compute warps' path
    const int s = (int)(n % kS);
    const uint32_t par = (n / kS) & 1;
    uint32_t* const buf = bufs + (size_t)s * kTile;
    mbar_wait(&ctl.full[s], par);
    const StageInfo si = ctl.info[s];
    const bool valid = si.tile >= 0;

    // ---- load 16 rows: warp w owns rows [w*512, w*512+512), segment j = 128 rows, lane-striped ----
    uint32_t v[kVecPerThread][4];
    uint32_t cnt = 0;  // four 8-bit counters: matches of this lane in segment j at bits 8j..8j+7
    const uint32_t e0 = warp * 512 + lane * 4;  // + j*128 + e
    if (valid) {
      if ((si.tma & 1) && si.len == kTile && (!kNullable || (si.row0 & 31) == 0)) {
        uint32_t vw[kVecPerThread];  // validity of this lane's rows: a nibble of one word per segment
        if (kNullable) {
          if (si.tma & 2) {
            const uint32_t* wp = vbufs + (size_t)s * (Cfg::kValidBytes / 4) + warp * 16 + (lane >> 3);
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) vw[j] = wp[j * 4] >> ((lane & 7) * 4);
          } else {
            const uint32_t* __restrict__ wp = a.valid + (si.row0 >> 5) + warp * 16 + (lane >> 3);
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) vw[j] = __ldg(wp + j * 4) >> ((lane & 7) * 4);
          }
        }
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
          const uint4 q = *reinterpret_cast<const uint4*>(buf + e0 + j * 128);
          v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
        }
        if (kNullable) {
#pragma unroll
          for (int j = 0; j < kVecPerThread; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!((vw[j] >> e) & 1u)) v[j][e] = never_lt<kCmp>();  // a null row never matches
        }
        uint32_t cnt_hi = 0;  // two chains halve the dependent-add latency
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) count_lt<kCmp>(j < 2 ? cnt : cnt_hi, v[j][e], a.thr, 1u << (8 * j));
        cnt += cnt_hi;
      } else {
        const uint32_t* __restrict__ src = a.in + si.row0;
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t i = e0 + j * 128 + e;
            // rows past the end of the tile never match
            v[j][e] = never_lt<kCmp>();
            if ((int32_t)i < si.len) {
              v[j][e] = (si.tma & 1) ? buf[i] : ld_stream_u32(src + i);
              if (kNullable) {
                const int64_t r = si.row0 + i;
                if (!((__ldg(a.valid + (r >> 5)) >> (r & 31)) & 1u)) v[j][e] = never_lt<kCmp>();
              }
            }
            cnt += ((int32_t)i < si.len ? lt_32<kCmp>(v[j][e], a.thr) : 0u) << (8 * j);
          }
        }
      }
    }
    // ---- ranks inside the warp: one scan of the packed counters covers all four segments ----
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - cnt;                           // packed, per segment
    const uint32_t seg = __shfl_sync(0xffffffffu, incl, 31);    // packed segment totals (<= 128 each)
    const uint32_t b1 = seg & 0xffu, b2 = b1 + ((seg >> 8) & 0xffu), b3 = b2 + ((seg >> 16) & 0xffu);
    const uint32_t wtotal = b3 + (seg >> 24);
    if (lane == 0) ctl.wtot[n & 1][warp] = wtotal;

    named_bar_sync(1, kCT);  // (A) the only CTA-wide barrier per tile (compute warps only)

    // ---- warp base and tile total ----
    uint32_t wv = lane < kW ? ctl.wtot[n & 1][lane] : 0u;
    uint32_t winc = wv;
#pragma unroll
    for (int o = 1; o < kW; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, winc, kW - 1);
    const uint32_t wbase = __shfl_sync(0xffffffffu, winc - wv, warp);
    if (valid && tid == 0) {
      // publish the count at all three levels as early as possible (fire and forget)
      st_relaxed_gpu_u32(a.agg + si.tile, kAggFlag | total);
      red_add_relaxed_gpu_u64(a.grp + (si.tile >> kGroupShift), (1ull << 48) | total);
      red_add_relaxed_gpu_u64(a.sgrp + (si.tile >> kSuperShift), (1ull << 48) | total);
      ctl.total[s] = total;
      mbar_arrive(&ctl.tot[s]);
    }

    // ---- write out the tile compacted kLag iterations ago (its global offset is known by now) ----
    auto write_out = [&](uint32_t m) {
      const int ps = (int)(m % kS);
      if (Cfg::kNamedBars) named_bar_sync(2 + ps, kCT + 32);
      else mbar_wait(&ctl.pre[ps], (m / kS) & 1);
      const uint32_t* __restrict__ stg = bufs + (size_t)ps * kTile;
      uint32_t* __restrict__ dst = a.out + ctl.prefix[ps];
      const uint32_t cnt_m = ctl.total[ps];
      for (uint32_t i = tid; i < cnt_m; i += kCT) st_stream_u32(dst + i, stg[i]);
      fence_proxy_async();
      if (Cfg::kNamedBars) {
        named_bar_arrive(2 + kS + ps, kCT + 32);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.empty[ps]);
      }
    };
    if (n >= (uint32_t)Cfg::kLag) write_out(n - Cfg::kLag);
    if (!valid) {  // drain: tiles n-kLag+1 .. n-1 were compacted before barrier A
      for (int m = (int)n - Cfg::kLag + 1; m < (int)n; ++m)
        if (m >= 0) write_out((uint32_t)m);
      break;
    }

    // ---- compact this tile in place (every thread read its rows before barrier A) ----
    // Hand-scheduled: compare, predicated store, predicated pointer bump — three instructions per
    // row on a byte address (the compiler's version recomputed the address per row: six).
    const uint32_t buf_s = smem_addr(buf);
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const uint32_t segbase = j == 0 ? 0u : (j == 1 ? b1 : (j == 2 ? b2 : b3));
      uint32_t p = buf_s + 4u * (wbase + segbase + ((excl >> (8 * j)) & 0xffu));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (kCmp == kCmpS32) {
          asm volatile(
              "{\n\t.reg .pred q;\n\t"
              "setp.lt.s32 q, %1, %2;\n\t"
              "@q st.shared.u32 [%0], %1;\n\t"
              "@q add.u32 %0, %0, 4;\n\t}"
              : "+r"(p)
              : "r"(v[j][e]), "r"(a.thr)
              : "memory");
        } else if (kCmp == kCmpF32) {
          asm volatile(
              "{\n\t.reg .pred q;\n\t.reg .f32 fa, fb;\n\t"
              "mov.b32 fa, %1;\n\tmov.b32 fb, %2;\n\t"
              "setp.lt.f32 q, fa, fb;\n\t"
              "@q st.shared.u32 [%0], %1;\n\t"
              "@q add.u32 %0, %0, 4;\n\t}"
              : "+r"(p)
              : "r"(v[j][e]), "r"(a.thr)
              : "memory");
        } else {
          asm volatile(
              "{\n\t.reg .pred q;\n\t"
              "setp.lt.u32 q, %1, %2;\n\t"
              "@q st.shared.u32 [%0], %1;\n\t"
              "@q add.u32 %0, %0, 4;\n\t}"
              : "+r"(p)
              : "r"(v[j][e]), "r"(a.thr)
              : "memory");
        }
      }
    }
  }
}

// Per-tile geometry of a ragged column: tile -> (first row, rows). tile_first[b] = first tile of
// batch b (exclusive prefix of the per-batch tile counts), nbatches+1 entries.
__global__ void filter_geom_kernel(const int64_t* __restrict__ batch_off,
                                   const int64_t* __restrict__ tile_first, int64_t nbatches,
                                   int64_t ntiles, int tile_rows, int64_t* __restrict__ tile_row0,
                                   int32_t* __restrict__ tile_len) {
  const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  int64_t lo = 0, hi = nbatches;  // last b with tile_first[b] <= tile
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (tile_first[mid] <= tile) lo = mid; else hi = mid;
  }
  const int64_t k = tile - tile_first[lo];
  const int64_t row0 = batch_off[lo] + k * tile_rows;
  const int64_t rest = batch_off[lo + 1] - row0;
  tile_row0[tile] = row0;
  tile_len[tile] = (int32_t)(rest < tile_rows ? rest : tile_rows);
}

// batch_end[b] = inclusive count at the last tile of batch b (carried over empty batches).
__global__ void filter_batch_end_kernel(const uint64_t* __restrict__ incl, int64_t nbatches,
                                        int64_t tiles_per_batch,
                                        const int64_t* __restrict__ tile_first,
                                        const int64_t* __restrict__ carry_in,
                                        int64_t* __restrict__ batch_end,
                                        int64_t* __restrict__ total) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t carry = carry_in ? *carry_in : 0;
  if (b < nbatches) {
    const int64_t last = tile_first ? tile_first[b + 1] : (b + 1) * tiles_per_batch;  // exclusive
    batch_end[b] = last > 0 ? (int64_t)incl[last - 1] : carry;
  }
  if (b == 0 && total) {
    const int64_t ntiles = tile_first ? tile_first[nbatches] : nbatches * tiles_per_batch;
    *total = ntiles > 0 ? (int64_t)incl[ntiles - 1] : carry;
  }
}

// ---- variants -----------------------------------------------------------------------------
// All variants share kTileRows so that the workspace size does not depend on the variant.
using Cfg0 = FilterCfg<256, 3, 4, 1>;   // 4096-row tiles, 3 x 16 KB stages, 4 CTAs/SM
using Cfg1 = FilterCfg<256, 4, 3, 1>;
using Cfg2 = FilterCfg<256, 4, 3, 2>;
using Cfg3 = FilterCfg<256, 5, 2, 2>;
using Cfg4 = FilterCfg<256, 5, 2, 1>;
using Cfg5 = FilterCfg<256, 6, 2, 3>;
using Cfg6 = FilterCfg<256, 4, 3, 2, true>;   // Cfg2 with named barriers for "prefix posted" / "stage drained"
using Cfg7 = FilterCfg<256, 5, 2, 3, true>;
constexpr int kNumVariants = 8;
constexpr int kTileRows = Cfg0::kTile;

// Kernel shape: ctx->tune[B2_TUNE_FILTER_VARIANT], default 6 = <256 threads, 4 stages, 3 CTAs/SM, lag 2,
// named barriers>: best on B200 (profiles/r1_filter.md). Every shape computes the same result.
#ifdef B2_LAB
int g_filter_debug = 0;  // tools/filter_lab.py (lab build, -DB2_LAB) sets this through b200olap_lab_filter_debug()
#endif

static inline int64_t tiles_of(int64_t len) { return (len + kTileRows - 1) / kTileRows; }

template <typename Cfg, bool kNullable = false, int kCmp = kCmpU32>
int launch_variant(b2_ctx* ctx, const FilterArgs& a, cudaStream_t s) {
  constexpr int kSmem = kNullable ? Cfg::kSmemBytesNullable : Cfg::kSmemBytes;
  static const int seen = b2_new_site();  // attribute + grid size are remembered per ctx (= per device)
  if (b2_first_use_on_device(ctx, seen)) {
    B2_CUDA_OK(ctx, cudaFuncSetAttribute(filter_lt_u32_kernel<Cfg, kNullable, kCmp>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int per_sm = 0;
    B2_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &per_sm, filter_lt_u32_kernel<Cfg, kNullable, kCmp>, Cfg::kThreads, kSmem));
    if (per_sm < 1) return b2_set_error(ctx, B2_ERR_CUDA, "filter kernel", "does not fit an SM");
    if (per_sm > Cfg::kCtasPerSm) per_sm = Cfg::kCtasPerSm;
    ctx->site_value[(size_t)seen] = per_sm * ctx->sm_count;
  }
  const int max_ctas = ctx->site_value[(size_t)seen];
  const int64_t grid = a.ntiles < max_ctas ? a.ntiles : max_ctas;
  filter_lt_u32_kernel<Cfg, kNullable, kCmp><<<(unsigned)grid, Cfg::kThreads, kSmem, s>>>(a);
  B2_LAUNCH_CHECK(ctx, "filter_lt_u32_kernel");
  return B2_OK;
}

// Workspace layout: header | agg | grp | sgrp  (all zeroed per launch)  | incl | ragged tables.
struct WsLayout {
  int64_t ntiles;
  size_t agg_off, grp_off, sgrp_off, zero_bytes, incl_off, tf_off, row0_off, len_off, total;
};
WsLayout ws_layout(int64_t ntiles, int64_t ragged_nbatches /* < 0: uniform */) {
  WsLayout w{};
  w.ntiles = ntiles;
  size_t o = sizeof(FilterWs);
  w.agg_off = o;   o += b2_align_up((size_t)ntiles * 4, 256);
  w.grp_off = o;   o += b2_align_up((size_t)((ntiles >> kGroupShift) + 1) * 8, 256);
  w.sgrp_off = o;  o += b2_align_up((size_t)((ntiles >> kSuperShift) + 1) * 8, 256);
  w.zero_bytes = o;
  w.incl_off = o;  o += b2_align_up((size_t)ntiles * 8, 256);
  if (ragged_nbatches >= 0) {
    w.tf_off = o;    o += b2_align_up((size_t)(ragged_nbatches + 1) * 8, 256);
    w.row0_off = o;  o += b2_align_up((size_t)ntiles * 8, 256);
    w.len_off = o;   o += b2_align_up((size_t)ntiles * 4, 256);
  }
  w.total = o;
  return w;
}
int64_t ragged_tiles(const int64_t* h_batch_off, int64_t nbatches) {
  int64_t n = 0;
  for (int64_t b = 0; b < nbatches; ++b) n += tiles_of(h_batch_off[b + 1] - h_batch_off[b]);
  return n;
}

int filter_launch(b2_ctx* ctx, int cmp, const uint32_t* d_in, const uint32_t* d_valid, int64_t nbatches, int64_t batch_len,
                  const int64_t* h_batch_off, const int64_t* d_batch_off, uint32_t thr,
                  uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                  const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, cudaStream_t s) {
  const bool uniform = (h_batch_off == nullptr);
  const int64_t ntiles = uniform ? nbatches * tiles_of(batch_len) : ragged_tiles(h_batch_off, nbatches);
  const WsLayout w = ws_layout(ntiles, uniform ? -1 : nbatches);
  if (ws_bytes < w.total || d_ws == nullptr)
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "filter workspace", "use b2_filter_ws_bytes()");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0, "workspace must be 16 B aligned");
  char* base = static_cast<char*>(d_ws);
  uint64_t* incl = reinterpret_cast<uint64_t*>(base + w.incl_off);
  int64_t* tile_first = nullptr;
  int64_t* tile_row0 = nullptr;
  int32_t* tile_len = nullptr;
  if (!uniform) {
    tile_first = reinterpret_cast<int64_t*>(base + w.tf_off);
    tile_row0 = reinterpret_cast<int64_t*>(base + w.row0_off);
    tile_len = reinterpret_cast<int32_t*>(base + w.len_off);
  }

  // zero the ticket and the count words of all three levels
  B2_CUDA_OK(ctx, cudaMemsetAsync(base, 0, w.zero_bytes, s));
  if (!uniform) {
    // tile_first is tiny; build it on the host and ship it (pageable copy is staged before return)
    std::string buf((size_t)(nbatches + 1) * 8, '\0');
    int64_t* tf = reinterpret_cast<int64_t*>(&buf[0]);
    tf[0] = 0;
    for (int64_t b = 0; b < nbatches; ++b) tf[b + 1] = tf[b] + tiles_of(h_batch_off[b + 1] - h_batch_off[b]);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(tile_first, tf, (size_t)(nbatches + 1) * 8, cudaMemcpyHostToDevice, s));
    if (ntiles > 0) {
      filter_geom_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(
          d_batch_off, tile_first, nbatches, ntiles, kTileRows, tile_row0, tile_len);
      B2_LAUNCH_CHECK(ctx, "filter_geom_kernel");
    }
  }
  if (ntiles > 0) {
    FilterArgs a{};
    a.in = d_in;
    a.valid = d_valid;
    a.out = d_out;
    a.thr = thr;
    a.ntiles = ntiles;
    a.batch_len = batch_len;
    a.tiles_per_batch = uniform ? tiles_of(batch_len) : 0;
    a.tile_row0 = tile_row0;
    a.tile_len = tile_len;
    a.carry_in = d_carry_in;
    a.ws = reinterpret_cast<FilterWs*>(base);
    a.agg = reinterpret_cast<uint32_t*>(base + w.agg_off);
    a.grp = reinterpret_cast<uint64_t*>(base + w.grp_off);
    a.sgrp = reinterpret_cast<uint64_t*>(base + w.sgrp_off);
    a.incl = incl;
#ifdef B2_LAB
    a.debug = g_filter_debug;
#endif
    if (cmp == kCmpS32) {  // typed and nullable kernels exist in one shape, the default one
      if (d_valid) B2_RETURN_NOT_OK((launch_variant<Cfg6, true, kCmpS32>(ctx, a, s)));
      else B2_RETURN_NOT_OK((launch_variant<Cfg6, false, kCmpS32>(ctx, a, s)));
    } else if (cmp == kCmpF32) {
      if (d_valid) B2_RETURN_NOT_OK((launch_variant<Cfg6, true, kCmpF32>(ctx, a, s)));
      else B2_RETURN_NOT_OK((launch_variant<Cfg6, false, kCmpF32>(ctx, a, s)));
    } else if (d_valid) {
      B2_RETURN_NOT_OK((launch_variant<Cfg6, true>(ctx, a, s)));
    } else switch (ctx->tune[B2_TUNE_FILTER_VARIANT]) {
      case 1: B2_RETURN_NOT_OK(launch_variant<Cfg1>(ctx, a, s)); break;
      case 2: B2_RETURN_NOT_OK(launch_variant<Cfg2>(ctx, a, s)); break;
      case 3: B2_RETURN_NOT_OK(launch_variant<Cfg3>(ctx, a, s)); break;
      case 4: B2_RETURN_NOT_OK(launch_variant<Cfg4>(ctx, a, s)); break;
      case 5: B2_RETURN_NOT_OK(launch_variant<Cfg5>(ctx, a, s)); break;
      case 6: B2_RETURN_NOT_OK(launch_variant<Cfg6>(ctx, a, s)); break;
      case 7: B2_RETURN_NOT_OK(launch_variant<Cfg7>(ctx, a, s)); break;
      default: B2_RETURN_NOT_OK(launch_variant<Cfg0>(ctx, a, s)); break;
    }
  }
  if (nbatches > 0 || d_total) {
    const int64_t nb = nbatches > 0 ? nbatches : 1;
    filter_batch_end_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(
        incl, nbatches, tiles_of(batch_len), tile_first, d_carry_in, d_batch_end, d_total);
    B2_LAUNCH_CHECK(ctx, "filter_batch_end_kernel");
  }
  return B2_OK;
}

}  // namespace

extern "C" {

#ifdef B2_LAB
// Lab build only (python -m dpu_olap_b200.build --lab): switches that make the kernel compute WRONG
// offsets on purpose, to time its parts in isolation. Not compiled into the product library.
int b200olap_lab_filter_debug(int bits) {
  g_filter_debug = bits;
  return B2_OK;
}
#endif

size_t b2_filter_ws_bytes(int64_t nbatches, int64_t batch_len) {
  if (nbatches < 0 || batch_len < 0) return 0;
  return ws_layout(nbatches * tiles_of(batch_len), -1).total;
}

size_t b2_filter_ragged_ws_bytes(const int64_t* h_batch_off, int64_t nbatches) {
  if (nbatches < 0 || (nbatches > 0 && !h_batch_off)) return 0;
  return ws_layout(ragged_tiles(h_batch_off, nbatches), nbatches).total;
}

int b2_filter_lt_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t nbatches, int64_t batch_len,
                         uint32_t threshold, uint32_t* d_out, int64_t* d_batch_end,
                         int64_t* d_total, const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                         void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches >= 0 && batch_len >= 0, "negative size");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  B2_REQUIRE(ctx, nbatches * batch_len == 0 || (d_in && d_out), "null column pointer");
  return filter_launch(ctx, kCmpU32, d_in, nullptr, nbatches, batch_len, nullptr, nullptr, threshold, d_out,
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

int b2_filter_lt_32_dev(b2_ctx* ctx, const void* d_in, int dtype, uint32_t threshold_bits,
                        const uint8_t* d_valid, int64_t nbatches, int64_t batch_len, void* d_out,
                        int64_t* d_batch_end, int64_t* d_total, const int64_t* d_carry_in, void* d_ws,
                        size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, dtype == B2_U32 || dtype == B2_I32 || dtype == B2_F32, "dtype must be B2_U32, B2_I32 or B2_F32");
  B2_REQUIRE(ctx, nbatches >= 0 && batch_len >= 0, "negative size");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  B2_REQUIRE(ctx, nbatches * batch_len == 0 || (d_in && d_out), "null column pointer");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_valid) & 3) == 0, "validity bitmap must be 4-byte aligned");
  return filter_launch(ctx, dtype, static_cast<const uint32_t*>(d_in), reinterpret_cast<const uint32_t*>(d_valid),
                       nbatches, batch_len, nullptr, nullptr, threshold_bits, static_cast<uint32_t*>(d_out),
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int b2_filter_lt_u32_nullable_dev(b2_ctx* ctx, const uint32_t* d_in, const uint8_t* d_valid,
                                  int64_t nbatches, int64_t batch_len, uint32_t threshold,
                                  uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                                  const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches >= 0 && batch_len >= 0, "negative size");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  B2_REQUIRE(ctx, nbatches * batch_len == 0 || (d_in && d_out), "null column pointer");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_valid) & 3) == 0, "validity bitmap must be 4-byte aligned");
  return filter_launch(ctx, kCmpU32, d_in, reinterpret_cast<const uint32_t*>(d_valid), nbatches, batch_len, nullptr,
                       nullptr, threshold, d_out, d_batch_end, d_total, d_carry_in, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

int b2_filter_lt_u32_ragged_dev(b2_ctx* ctx, const uint32_t* d_in, const int64_t* h_batch_off,
                                const int64_t* d_batch_off, int64_t nbatches, uint32_t threshold,
                                uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                                const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                                void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches >= 0, "negative size");
  B2_REQUIRE(ctx, h_batch_off && d_batch_off, "batch offset tables are null");
  for (int64_t b = 0; b < nbatches; ++b)
    B2_REQUIRE(ctx, h_batch_off[b + 1] >= h_batch_off[b], "batch offsets must be non-decreasing");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  return filter_launch(ctx, kCmpU32, d_in, nullptr, nbatches, 0, h_batch_off, d_batch_off, threshold, d_out,
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

int b2_filter_lt_32_ragged_dev(b2_ctx* ctx, const void* d_in, int dtype, uint32_t threshold_bits,
                               const uint8_t* d_valid, const int64_t* h_batch_off, const int64_t* d_batch_off,
                               int64_t nbatches, void* d_out, int64_t* d_batch_end, int64_t* d_total,
                               const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, dtype == B2_U32 || dtype == B2_I32 || dtype == B2_F32, "dtype must be B2_U32, B2_I32 or B2_F32");
  B2_REQUIRE(ctx, nbatches >= 0, "negative size");
  B2_REQUIRE(ctx, h_batch_off && d_batch_off, "batch offset tables are null");
  for (int64_t b = 0; b < nbatches; ++b)
    B2_REQUIRE(ctx, h_batch_off[b + 1] >= h_batch_off[b], "batch offsets must be non-decreasing");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_valid) & 3) == 0, "validity bitmap must be 4-byte aligned");
  return filter_launch(ctx, dtype, static_cast<const uint32_t*>(d_in), reinterpret_cast<const uint32_t*>(d_valid),
                       nbatches, 0, h_batch_off, d_batch_off, threshold_bits, static_cast<uint32_t*>(d_out),
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
