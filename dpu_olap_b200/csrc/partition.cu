// partition.cu — radix hash partitioning of (key, value) rows.
//
// Replaces the reference's DPU partition program (dpu/shared/kernels/partition.c:296-341):
// build_histogram (:67-92, one mutex-protected increment per row), prefix_sum (:94-137) and
// partition_array / write_new_output (:167-294, a mutex per row, 8-byte MRAM read-modify-write),
// plus the host-side offset bookkeeping (Partitioner::GetOffsets, partitioner.cc:280-312) and the
// scatter-gather DMA that brings the pieces of a partition together (LoadPartitions, :350-375).
// The hash and bucket are the reference's: wang_hash_uint32 (partition.c:20-28) and the top bits
// of the hash (BUCKET_OF, partition.c:45-46).
//
// B200 design — three kernels per pass, every row read twice (once keys-only) and written once:
//   1. part_hist_kernel     per work unit (a run of 8192-row tiles) a histogram over the 2^bits
//                           buckets with shared-memory atomics (measured on B200: 1.3e12
//                           lane-ops/s, an order of magnitude above a __match_any_sync scheme),
//                           eight independent key loads in flight per thread.
//   2. exclusive scan       the histogram is laid out (segment, bucket, unit)-major, so one flat
//                           scan (scan.cu, decoupled look-back) yields the global destination of
//                           every (bucket, unit) run — no per-partition mutex, no host round trip.
//   3. scatter              per tile: rank rows inside their bucket (atomicAdd on the tile's
//                           shared-memory counters returns the rank), scan the 2^bits tile counts,
//                           stage the tile SORTED BY BUCKET in shared memory, then write every
//                           bucket's run out. The tile after the current one is requested into L2 by
//                           the copy engine (cp.async.bulk.prefetch.L2) at the top of the tile, so the
//                           register loads issued once the tile is staged hit L2 and the DRAM reads
//                           run under the rank / scan / stage steps.
//      part_scatter_bulk_kernel     fan-outs from 2^8 (the join's passes): rows that do not complete a
//                           32-byte sector are held back per bucket (B200's L2 fills a partially
//                           written sector from DRAM on a write miss) and a bucket's whole sectors
//                           leave as ONE shared -> global bulk copy per tile (UBLKCP): the stores are
//                           the copy engine's, asynchronous to the SM. 512 threads x 2 CTAs/SM up to
//                           2^9 buckets, 1024 threads x 1 CTA/SM at 2^10; peer mode for the shuffle.
//      part_scatter_kernel          small fan-outs: consecutive threads write consecutive rows of a run.
//      part_scatter_lines_kernel    the multi-GPU shuffle: whole 128-byte lines to peer memory over
//                           NVLink, optionally with a CTA budget (persistent walk over the work units).
//      part_scatter_sectors_kernel, part_scatter_quads_kernel: earlier whole-sector kernels with the
//                           flush done by the threads; selectable (B2_TUNE_SCATTER_SECTOR_TILE) and
//                           parity-tested, slower than the bulk kernel (profiles/r2_scatter_bulk.md).
// A pass can be "segmented": each input segment (= a partition of the previous pass) is
// partitioned independently, which is how the join refines 2^10 coarse partitions into up to
// 2^20 shared-memory-sized ones while every pass keeps >= 64 B write runs.
// Rows are carried as 8-byte (key, value) pairs between passes (one 8 B access per row instead
// of two 4 B accesses to two columns).
#include "partition.cuh"

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "scan.cuh"
#include "tma.cuh"

namespace {

constexpr int kThreads = kPartThreads;

template <bool kAoS>
__device__ __forceinline__ void load_row(const PartInput& in, int64_t row, uint32_t& k, uint32_t& v) {
  if (kAoS) {
    const uint2 p = ld_stream_v2(in.pairs + row);
    k = p.x;
    v = p.y;
  } else {
    k = ld_stream_u32(in.keys + row);
    v = in.vals ? ld_stream_u32(in.vals + row) : (uint32_t)row;
  }
}
template <bool kAoS>
__device__ __forceinline__ uint32_t load_key(const PartInput& in, int64_t row) {
  return kAoS ? ld_stream_u32(reinterpret_cast<const uint32_t*>(in.pairs + row))
              : ld_stream_u32(in.keys + row);
}

struct Unit {
  int64_t row0, row1;  // rows of this unit
  int64_t hbase;       // histogram entry of (bucket 0, this unit); bucket p is at hbase + p*ustride
  int64_t ustride;     // units in this unit's segment
  int64_t hbase0;      // histogram entry of (bucket 0, first unit of the segment)
  bool valid;
};

// Work unit u -> (segment, k-th unit of the segment). unit_first[s] = number of units in the
// segments before s; empty segments own no unit.
__device__ __forceinline__ Unit find_unit(const int64_t* __restrict__ seg_off,
                                          const int64_t* __restrict__ unit_first, int64_t nseg,
                                          int64_t unit_rows, int P, int64_t id);
__device__ __forceinline__ Unit find_unit(const int64_t* __restrict__ seg_off,
                                          const int64_t* __restrict__ unit_first, int64_t nseg,
                                          int64_t unit_rows, int P) {
  return find_unit(seg_off, unit_first, nseg, unit_rows, P, (int64_t)blockIdx.x);
}
__device__ __forceinline__ Unit find_unit(const int64_t* __restrict__ seg_off,
                                          const int64_t* __restrict__ unit_first, int64_t nseg,
                                          int64_t unit_rows, int P, int64_t id) {
  Unit u;
  u.valid = id < unit_first[nseg];
  if (!u.valid) return u;
  int64_t lo = 0, hi = nseg;  // last s with unit_first[s] <= id
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (unit_first[mid] <= id) lo = mid; else hi = mid;
  }
  const int64_t k = id - unit_first[lo];
  u.ustride = unit_first[lo + 1] - unit_first[lo];
  u.row0 = seg_off[lo] + k * unit_rows;
  u.row1 = min(u.row0 + unit_rows, seg_off[lo + 1]);
  u.hbase0 = unit_first[lo] * P;
  u.hbase = u.hbase0 + k;
  return u;
}

// unit_first[s] for all segments (one CTA; nseg <= a few thousand).
__global__ void part_unit_table_kernel(const int64_t* __restrict__ seg_off, int64_t nseg,
                                       int64_t unit_rows, int64_t* __restrict__ unit_first) {
  __shared__ int64_t carry;
  __shared__ int64_t warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < nseg; base += blockDim.x) {
    const int64_t s = base + threadIdx.x;
    int64_t c = 0;
    if (s < nseg) c = (seg_off[s + 1] - seg_off[s] + unit_rows - 1) / unit_rows;
    int64_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
      int64_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const int64_t excl = carry + warp_tot[warp] + incl - c;
    if (s < nseg) unit_first[s] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) unit_first[nseg] = carry;
}

// Hash + selection + bucket of one key; returns 0xffffffff for rows outside the hash-space slice.
// kValPred (compile time: the plain passes must not pay for it): also drop rows whose value fails
// the pushed-down predicate val < g.val_thr.
// The hash-space slice test as ONE mask-and-compare per row: selected <=> (h & mask) == cmp, with
// mask = cmp = 0 when every row is taken (part_selected() spelled out costs two shifts more).
struct SliceSel {
  uint32_t mask, cmp;
};
__device__ __forceinline__ SliceSel slice_sel(const PartGeom& g) {
  SliceSel s{0u, 0u};
  if (g.sel_bits > 0) {
    const int sh = 32 - g.sel_shl - g.sel_bits;
    s.mask = ((1u << g.sel_bits) - 1u) << sh;
    s.cmp = g.sel_val << sh;
  }
  return s;
}
template <bool kValPred>
__device__ __forceinline__ uint32_t bucket_or_skip(uint32_t key, uint32_t val, const PartGeom& g,
                                                   const SliceSel& sel) {
  const uint32_t h = wang_hash_u32(key);
  const uint32_t b = part_bucket(h, g.shl, g.bits);
  const bool keep = (h & sel.mask) == sel.cmp && (!kValPred || val < g.val_thr);
  return keep ? b : 0xffffffffu;
}

// The same with the value predicate decided at run time (the peer scatter kernel has one shape).
__device__ __forceinline__ uint32_t bucket_or_skip_rt(uint32_t key, uint32_t val, const PartGeom& g,
                                                      const SliceSel& sel) {
  const uint32_t h = wang_hash_u32(key);
  const uint32_t b = part_bucket(h, g.shl, g.bits);
  const bool keep = (h & sel.mask) == sel.cmp && (!g.val_pred || val < g.val_thr);
  return keep ? b : 0xffffffffu;
}

template <bool kAoS, bool kValPred>
__global__ void __launch_bounds__(kThreads, 4)
part_hist_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                 const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                 PartGeom g, uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[1 << kPartMaxBits];
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P);
  if (!u.valid) return;
  const uint32_t tid = threadIdx.x;
  for (int i = tid; i < P; i += kThreads) cnt[i] = 0;
  __syncthreads();
  constexpr int kU = 8;  // independent loads in flight per thread
  int64_t base = u.row0;
  if (!kAoS && !kValPred && ((reinterpret_cast<uintptr_t>(in.keys + u.row0) & 15) == 0)) {
    // key column, 16-byte aligned: four keys per load (a quarter of the load and address
    // instructions of the per-key loop; the histogram of a key column is instruction-bound)
    const uint4* __restrict__ k4 = reinterpret_cast<const uint4*>(in.keys + u.row0);
    const int64_t nvec = (u.row1 - u.row0) >> 2;
    int64_t v = 0;
    for (; v + (int64_t)kThreads * 4 <= nvec; v += (int64_t)kThreads * 4) {
      uint4 kk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) kk[q] = ld_stream_v4(k4 + v + q * kThreads + tid);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {kk[q].x, kk[q].y, kk[q].z, kk[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (sel.mask == 0) {  // every row is kept: no per-row branch (uniform test)
            atomicAdd(&cnt[part_bucket(wang_hash_u32(w[e]), g.shl, g.bits)], 1u);
          } else {
            const uint32_t b = bucket_or_skip<false>(w[e], 0u, g, sel);
            if (b != 0xffffffffu) atomicAdd(&cnt[b], 1u);
          }
        }
      }
    }
    base = u.row0 + (v << 2);  // the rest goes through the per-key loops below
  }
  // full blocks: no per-row bounds checks
  for (; base + (int64_t)kThreads * kU <= u.row1; base += (int64_t)kThreads * kU) {
    uint32_t key[kU], val[kU];
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      val[q] = 0;
      if (kValPred) load_row<kAoS>(in, base + q * kThreads + tid, key[q], val[q]);  // the predicate needs the value
      else key[q] = load_key<kAoS>(in, base + q * kThreads + tid);
    }
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      if (!kValPred && sel.mask == 0) {
        atomicAdd(&cnt[part_bucket(wang_hash_u32(key[q]), g.shl, g.bits)], 1u);
      } else {
        const uint32_t b = bucket_or_skip<kValPred>(key[q], val[q], g, sel);
        if (b != 0xffffffffu) atomicAdd(&cnt[b], 1u);
      }
    }
  }
  for (int64_t row = base + tid; row < u.row1; row += kThreads) {
    uint32_t key, val = 0;
    if (kValPred) load_row<kAoS>(in, row, key, val);
    else key = load_key<kAoS>(in, row);
    const uint32_t b = bucket_or_skip<kValPred>(key, val, g, sel);
    if (b != 0xffffffffu) atomicAdd(&cnt[b], 1u);
  }
  __syncthreads();
  for (int p = tid; p < P; p += kThreads) hist[u.hbase + (int64_t)p * u.ustride] = cnt[p];
}

// Scatter. Per 8192-row tile: every row is hashed ONCE; its rank inside its bucket comes from a
// shared-memory atomicAdd; after a scan of the tile's bucket counts the rows are staged in shared
// memory sorted by bucket TOGETHER with their 16-bit bucket id, so the stream-out needs one
// 8-byte table read (destination of the bucket's run minus its position in the tile) per row
// instead of hashing the key again. The first version of this kernel spent 105-111
// lane-instructions per row (profiles/r1_join.md).
template <bool kAoS, int kT, int kI, int kCtas, bool kValPred, bool kPre>
__global__ void __launch_bounds__(kT, kCtas)
part_scatter_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                    const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                    PartGeom g, const uint64_t* __restrict__ scanned, uint2* __restrict__ out,
                    int64_t out_cap, const uint64_t* __restrict__ bucket_addr,
                    unsigned int* __restrict__ overflow) {
  // kT threads handle tiles of kT * kI rows, kCtas CTAs per SM (shape chosen by the launcher).
  constexpr int kTileRows = kT * kI;
  constexpr int kW = kT / 32;
  constexpr int kBpt = (1 << kPartMaxBits) / kT;  // buckets per thread in the tile scan
  extern __shared__ __align__(16) unsigned char smem[];
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P);
  if (!u.valid) return;
  // shared-memory carve-up
  uint2* stage = reinterpret_cast<uint2*>(smem);                                   // [kTileRows]
  uint64_t* gbase = reinterpret_cast<uint64_t*>(smem + sizeof(uint2) * kTileRows);  // [P] next BYTE address
  uint64_t* delta = gbase + P;                                                      // [P] gbase - 8*tile_start
  uint32_t* tile_start = reinterpret_cast<uint32_t*>(delta + P);                    // [P]
  uint32_t* tile_cnt = tile_start + P;                                              // [P]
  uint16_t* sbucket = reinterpret_cast<uint16_t*>(tile_cnt + P);                    // [kTileRows]
  __shared__ uint32_t warp_tot[kW];
  __shared__ uint32_t s_tile_total;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Destinations are byte addresses. Local mode: out + 8 * (flat scan position). Peer mode
  // (bucket_addr != nullptr, the fused multi-GPU shuffle): bucket p of THIS rank starts at
  // bucket_addr[p] — an address inside another GPU's receive buffer, written over NVLink — and
  // this unit's run starts where the units before it in the segment end.
  const uint64_t cap_addr = bucket_addr ? ~0ull : reinterpret_cast<uint64_t>(out) + 8ull * (uint64_t)out_cap;
  for (int p = tid; p < P; p += kT) {
    const uint64_t pos = scanned[u.hbase + (int64_t)p * u.ustride];
    gbase[p] = bucket_addr ? bucket_addr[p] + 8ull * (pos - scanned[u.hbase0 + (int64_t)p * u.ustride])
                           : reinterpret_cast<uint64_t>(out) + 8ull * pos;
    tile_cnt[p] = 0;
  }
  __syncthreads();

  // kPre: the rows of the NEXT tile are requested right after the current tile is staged (their
  // registers are free from then on), so the loads fly during the stream-out instead of stalling
  // the next rank step.
  uint32_t key[kI], val[kI];
  auto load_tile = [&](int64_t t0) {
    if (t0 + kTileRows <= u.row1) {  // full tile: no bounds checks
#pragma unroll
      for (int it = 0; it < kI; ++it) load_row<kAoS>(in, t0 + it * kT + tid, key[it], val[it]);
    } else {
#pragma unroll
      for (int it = 0; it < kI; ++it) {
        const int64_t row = t0 + it * kT + tid;
        key[it] = 0;
        val[it] = 0;
        if (row < u.row1) load_row<kAoS>(in, row, key[it], val[it]);
      }
    }
  };
  if (kPre && u.row0 < u.row1) load_tile(u.row0);

  for (int64_t t0 = u.row0; t0 < u.row1; t0 += kTileRows) {
    if (kPre && tid == 0 && t0 + kTileRows < u.row1) {  // the next tile goes to L2 now, to registers after this tile is staged
      const int64_t n0 = t0 + kTileRows, nn = min((int64_t)kTileRows, u.row1 - n0);
      if (kAoS) {
        l2_prefetch(in.pairs + n0, nn * 8);
      } else {
        l2_prefetch(in.keys + n0, nn * 4);
        if (in.vals) l2_prefetch(in.vals + n0, nn * 4);
      }
    }
    // ---- (load,) hash once, rank inside the bucket ----
    if (!kPre) load_tile(t0);
    uint32_t packed[kI];  // bucket | rank << 16
    if (t0 + kTileRows <= u.row1) {
      if (!kValPred && sel.mask == 0) {
        // every row is kept (no slice, no predicate): no per-row branch around the ranking atomic
#pragma unroll
        for (int it = 0; it < kI; ++it) {
          const uint32_t b = part_bucket(wang_hash_u32(key[it]), g.shl, g.bits);
          packed[it] = b | (atomicAdd(&tile_cnt[b], 1u) << 16);  // rank < 8192
        }
      } else {
#pragma unroll
        for (int it = 0; it < kI; ++it) {
          const uint32_t b = bucket_or_skip<kValPred>(key[it], val[it], g, sel);
          packed[it] = b;
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&tile_cnt[b], 1u) << 16);  // rank < 8192
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < kI; ++it) {
        const int64_t row = t0 + it * kT + tid;
        packed[it] = 0xffffffffu;
        if (row < u.row1) {
          const uint32_t b = bucket_or_skip<kValPred>(key[it], val[it], g, sel);
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&tile_cnt[b], 1u) << 16);
        }
      }
    }
    __syncthreads();

    // ---- exclusive scan of the tile counts (kBpt buckets per thread) ----
    uint32_t c[kBpt];
    uint32_t csum = 0;
#pragma unroll
    for (int q = 0; q < kBpt; ++q) {
      const int p = kBpt * tid + q;
      c[q] = p < P ? tile_cnt[p] : 0;
      csum += c[q];
    }
    uint32_t incl = csum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    {
      // every warp scans the 16 warp totals itself (no second barrier for a one-warp scan)
      const uint32_t w = lane < kW ? warp_tot[lane] : 0;
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      const uint32_t wbase = __shfl_sync(0xffffffffu, wi - w, warp);
      const uint32_t tile_total = __shfl_sync(0xffffffffu, wi, kW - 1);
      if (tid == 0) s_tile_total = tile_total;
      uint32_t excl = wbase + incl - csum;
#pragma unroll
      for (int q = 0; q < kBpt; ++q) {
        const int p = kBpt * tid + q;
        if (p < P) {
          tile_start[p] = excl;
          delta[p] = gbase[p] - 8ull * excl;
        }
        excl += c[q];
      }
    }
    __syncthreads();

    // ---- stage the tile sorted by bucket, bucket id beside the row ----
#pragma unroll
    for (int it = 0; it < kI; ++it) {
      if (packed[it] != 0xffffffffu) {
        const uint32_t b = packed[it] & 0xffffu;
        const uint32_t pos = tile_start[b] + (packed[it] >> 16);
        stage[pos] = make_uint2(key[it], val[it]);
        sbucket[pos] = (uint16_t)b;
      }
    }
    __syncthreads();

    if (kPre && t0 + kTileRows < u.row1) load_tile(t0 + kTileRows);

    // ---- stream out: consecutive threads -> consecutive rows of a bucket's run ----
    const uint32_t total = s_tile_total;
    for (uint32_t j = tid; j < total; j += kT) {
      const uint2 kv = stage[j];
      const uint64_t dst = delta[sbucket[j]] + 8ull * j;
      if (dst < cap_addr) {
        st_stream_v2(reinterpret_cast<uint2*>(dst), kv);
      } else if (overflow) {
        *overflow = 1u;
      }
    }
    __syncthreads();
    // ---- advance the running destinations, clear the counters ----
    for (int p = tid; p < P; p += kT) {
      gbase[p] += 8ull * tile_cnt[p];
      tile_cnt[p] = 0;
    }
    __syncthreads();
  }
}

// ---- peer-destination scatter with line carry --------------------------------------------------
// NVLink peer stores run at ~670 GB/s when every warp-level store covers whole 128-byte lines of
// the destination, but only ~380 GB/s for 128-byte runs that start 8 bytes off a line
// (tools/p2p_write_probe.py, profiles/r1_nvlink_store_probe.log). Rows are packed densely at the
// destination, so a bucket's run of one tile starts at an arbitrary 8-byte offset. This kernel
// therefore keeps, per bucket, the rows that do not fill a whole 128-byte line (at most 15) in a
// shared-memory carry buffer and prepends them to the bucket's rows of the next tile: only whole,
// aligned lines are stored (one half-warp per line), plus one partial line per (unit, bucket) at
// each end of the unit's region.
constexpr int kLineRows = 16;                      // 128 B / 8 B
constexpr int kLcThreads = 1024;
constexpr int kLcItems = 8;
constexpr int kLcTile = kLcThreads * kLcItems;     // 8192 rows
constexpr int kLcMaxLines = kLcTile / kLineRows + (1 << kPartMaxBits);  // whole lines one tile can flush

struct LcSmem {  // dynamic shared memory of part_scatter_lines_kernel (P = 1024 buckets)
  uint2 stage[kLcTile];                                  // 64 KB: the tile sorted by bucket
  uint2 carry[(1 << kPartMaxBits) * kLineRows];          // 128 KB: held-back rows, 16 slots per bucket
  uint64_t gline[1 << kPartMaxBits];                     // aligned byte address of carry slot 0
  uint32_t tile_start[1 << kPartMaxBits];
  uint32_t tile_cnt[1 << kPartMaxBits];
  uint32_t ccnt[1 << kPartMaxBits];                      // carried rows | ghost rows << 8
  uint32_t line_off[1 << kPartMaxBits];                  // first flushed line of the bucket in this tile
  uint16_t line_bucket[kLcMaxLines + 16];
  uint32_t warp_tot[kLcThreads / 32];
  uint32_t n_lines;
};

template <bool kAoS>
__global__ void __launch_bounds__(kLcThreads, 1)
part_scatter_lines_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                          const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                          PartGeom g, const uint64_t* __restrict__ scanned,
                          const uint64_t* __restrict__ bucket_addr, const int64_t* __restrict__ abort_flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  LcSmem& sm = *reinterpret_cast<LcSmem*>(smem);
  // the exchange was called off on the device (a receive buffer would overflow): store nothing
  if (abort_flag && *abort_flag) return;
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  constexpr int kW = kLcThreads / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // A grid smaller than the number of work units walks them (a CTA budget: the rest of the SMs stay
  // free for a kernel of another stream, e.g. the build side's fine pass under the probe side's scatter).
  for (int64_t unit_id = blockIdx.x;; unit_id += gridDim.x) {
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P, unit_id);
  if (!u.valid) return;
  __syncthreads();  // the previous unit's last reads of the carry buffers

  if ((int)tid < P) {
    const int p = tid;
    const uint64_t a0 = bucket_addr[p] +
                        8ull * (scanned[u.hbase + (int64_t)p * u.ustride] - scanned[u.hbase0 + (int64_t)p * u.ustride]);
    const uint32_t ghost = (uint32_t)((a0 >> 3) & (kLineRows - 1));  // rows of the line that are not ours
    sm.gline[p] = a0 - 8ull * ghost;
    sm.ccnt[p] = ghost | (ghost << 8);
    sm.tile_cnt[p] = 0;
  }
  __syncthreads();

  for (int64_t t0 = u.row0; t0 < u.row1; t0 += kLcTile) {
    if (tid == 0 && t0 + kLcTile < u.row1) {  // the next tile goes to L2 while this one is ranked, staged and sent
      const int64_t n0 = t0 + kLcTile, nn = min((int64_t)kLcTile, u.row1 - n0);
      if (kAoS) {
        l2_prefetch(in.pairs + n0, nn * 8);
      } else {
        l2_prefetch(in.keys + n0, nn * 4);
        if (in.vals) l2_prefetch(in.vals + n0, nn * 4);
      }
    }
    // ---- load, hash once, rank inside the bucket ----
    uint32_t key[kLcItems], val[kLcItems], packed[kLcItems];  // packed = bucket | rank << 16
#pragma unroll
    for (int it = 0; it < kLcItems; ++it) {
      const int64_t row = t0 + it * kLcThreads + tid;
      key[it] = 0;
      val[it] = 0;
      packed[it] = 0xffffffffu;
      if (row < u.row1) load_row<kAoS>(in, row, key[it], val[it]);
    }
#pragma unroll
    for (int it = 0; it < kLcItems; ++it) {
      const int64_t row = t0 + it * kLcThreads + tid;
      if (row < u.row1) {
        const uint32_t b = bucket_or_skip_rt(key[it], val[it], g, sel);  // a pushed-down predicate drops rows here
        if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
      }
    }
    __syncthreads();

    // ---- one scan for both the staged position and the first flushed line of every bucket ----
    uint32_t tc = 0, cc = 0, nl = 0;
    if ((int)tid < P) {
      tc = sm.tile_cnt[tid];
      cc = sm.ccnt[tid] & 0xffu;
      nl = (cc + tc) / kLineRows;
    }
    const uint32_t mine = tc | (nl << 16);  // rows <= 8192 and lines <= 1536 never carry into each other
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    {
      const uint32_t w = sm.warp_tot[lane];  // kW == 32 warps
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      const uint32_t wbase = __shfl_sync(0xffffffffu, wi - w, warp);
      const uint32_t all = __shfl_sync(0xffffffffu, wi, kW - 1);
      if (tid == 0) sm.n_lines = all >> 16;
      const uint32_t excl = wbase + incl - mine;
      if ((int)tid < P) {
        sm.tile_start[tid] = excl & 0xffffu;
        sm.line_off[tid] = excl >> 16;
        for (uint32_t li = 0; li < nl; ++li) sm.line_bucket[(excl >> 16) + li] = (uint16_t)tid;
      }
    }
    __syncthreads();

    // ---- stage the tile sorted by bucket ----
#pragma unroll
    for (int it = 0; it < kLcItems; ++it) {
      if (packed[it] != 0xffffffffu) {
        const uint32_t b = packed[it] & 0xffffu;
        sm.stage[sm.tile_start[b] + (packed[it] >> 16)] = make_uint2(key[it], val[it]);
      }
    }
    __syncthreads();

    // ---- flush whole lines: one half-warp per 128-byte line of the destination ----
    {
      const uint32_t n_lines = sm.n_lines;
      const uint32_t hw = tid >> 4, l = tid & 15;
      for (uint32_t i = hw; i < n_lines; i += kLcThreads / 16) {
        const uint32_t b = sm.line_bucket[i];
        const uint32_t li = i - sm.line_off[b];
        const uint32_t cw = sm.ccnt[b];
        const uint32_t c0 = cw & 0xffu, ghost = cw >> 8;
        const uint32_t e = li * kLineRows + l;  // position in (carried rows ++ staged rows)
        if (e >= ghost) {                       // ghost rows belong to the previous unit's region
          const uint2 kv = e < c0 ? sm.carry[b * kLineRows + e] : sm.stage[sm.tile_start[b] + e - c0];
          st_stream_v2(reinterpret_cast<uint2*>(sm.gline[b] + 8ull * e), kv);
        }
      }
    }
    __syncthreads();

    // ---- carry the rows that did not fill a line; advance; clear the counters ----
    if ((int)tid < P) {
      const uint32_t b = tid;
      const uint32_t tot = cc + tc;
      const uint32_t rem = tot & (kLineRows - 1);
      uint32_t ghost = sm.ccnt[b] >> 8;
      if (nl == 0) {  // nothing flushed: append the tile's rows to the carry
        for (uint32_t k = 0; k < tc; ++k) sm.carry[b * kLineRows + cc + k] = sm.stage[sm.tile_start[b] + k];
      } else {        // the tail comes from the staged rows (cc < 16 <= 16 * nl)
        const uint32_t from = sm.tile_start[b] + nl * kLineRows - cc;
        for (uint32_t k = 0; k < rem; ++k) sm.carry[b * kLineRows + k] = sm.stage[from + k];
        sm.gline[b] += 128ull * nl;
        ghost = 0;
      }
      sm.ccnt[b] = rem | (ghost << 8);
      sm.tile_cnt[b] = 0;
    }
    __syncthreads();
  }

  // ---- end of the unit: the last, partial line of every bucket ----
  {
    const uint32_t hw = tid >> 4, l = tid & 15;
    for (int b = hw; b < P; b += kLcThreads / 16) {
      const uint32_t cw = sm.ccnt[b];
      const uint32_t c0 = cw & 0xffu, ghost = cw >> 8;
      if (l >= ghost && l < c0)
        st_stream_v2(reinterpret_cast<uint2*>(sm.gline[b] + 8ull * l), sm.carry[b * kLineRows + l]);
    }
  }
  }  // units of this CTA
}

// ---- local scatter that only writes whole 32-byte sectors ------------------------------------------
// With a fan-out of 512-1024 a bucket's run in one 8192-row tile is 8-16 rows (64-128 B) and starts
// at an arbitrary 8-byte offset, so most runs begin and end with a PARTIAL 32-byte sector. ncu at
// 2^27 rows (profiles/r1_scatter_fanout.md): B200's L2 fills a partially written sector from DRAM
// as soon as the write misses (13.6 M fills, +40 % DRAM reads, 1.46 sectors written per sector of
// payload) and the LSU backs up behind those stores (lg_throttle 5.2 stall cycles per issue against
// 0.6 at a fan-out of 256) — the pass takes 1.8x as long per row. This kernel therefore holds back,
// per bucket, the rows that do not complete a sector (at most 3, in shared memory) and prepends them
// to the bucket's rows of the next tile; four adjacent lanes store one whole, aligned sector. Only
// the first and the last sector of a (unit, bucket) run can be partial.
constexpr int kScRows = 4;                       // rows per 32-byte sector

struct ScDesc {       // per bucket and tile, read once per flushed row
  int32_t sbase;      // stage index of (carried ++ staged) position 0, minus 4 * (first flushed sector)
  uint32_t dsec;      // destination sector of position 0, minus (first flushed sector); between two
                      // tiles: the next unwritten sector of the run, counted from the output base
};
// Shape of the kernel: kT threads x kI rows per thread per tile. <512, 16> runs two CTAs per SM
// (8192-row tiles: a bucket's run is 64 B at a fan-out of 1024); <1024, 16> runs one CTA per SM over
// 16384-row tiles, which doubles the run length (128 B: the memory system takes 128-byte runs at
// 5.2 TB/s against 4.0 TB/s for 64-byte ones, tools/microbench_scatter_ceiling.cu) and halves the
// per-bucket bookkeeping per row.
template <int kT, int kI>
struct ScShape {
  static constexpr int kTile = kT * kI;
  static constexpr int kBpt = (1 << kPartMaxBits) / kT;                              // buckets per thread
  static constexpr int kMaxSec = (kTile + 3 * (1 << kPartMaxBits)) / kScRows;        // whole sectors one tile can flush
  static constexpr int kCtas = kT >= 1024 ? 1 : 2;
};
template <int kT, int kI>
struct ScSmemT {
  uint2 stage[ScShape<kT, kI>::kTile];          // the tile's rows sorted by bucket
  uint2 carry[(1 << kPartMaxBits) * 3];         // 24 KB: held-back rows (slots 0..ghost-1 are not ours)
  ScDesc bd[1 << kPartMaxBits];                 // 8 KB
  uint32_t tile_cnt[1 << kPartMaxBits];         // carried + rows ranked so far
  int32_t sbase[1 << kPartMaxBits];             // tile_start - carried
  uint32_t meta[1 << kPartMaxBits];             // first flushed sector (16 bits) | carried << 16 | ghost << 18
  uint16_t sec_desc[ScShape<kT, kI>::kMaxSec + 16];  // flushed sector -> bucket | (first sector: carried << 10 | ghost << 12)
  uint32_t warp_tot[kT / 32];
  uint32_t n_sec;
};

template <bool kAoS, bool kPre, int kScT, int kScI>
__global__ void __launch_bounds__(kScT, (ScShape<kScT, kScI>::kCtas))
part_scatter_sectors_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                            const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                            PartGeom g, const uint64_t* __restrict__ scanned, uint2* __restrict__ out,
                            int64_t out_cap, unsigned int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem[];
  using Shape = ScShape<kScT, kScI>;
  constexpr int kScTile = Shape::kTile, kScBpt = Shape::kBpt;
  ScSmemT<kScT, kScI>& sm = *reinterpret_cast<ScSmemT<kScT, kScI>*>(smem);
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P);
  if (!u.valid) return;
  constexpr int kW = kScT / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t cap = (uint64_t)out_cap;  // rows

#pragma unroll
  for (int q = 0; q < kScBpt; ++q) {
    const int p = kScBpt * tid + q;
    if (p < P) {
      const uint64_t pos = scanned[u.hbase + (int64_t)p * u.ustride];  // first output row of (unit, bucket)
      const uint32_t ghost = (uint32_t)(pos & (kScRows - 1));         // rows of that sector that are not ours
      sm.bd[p].dsec = (uint32_t)(pos >> 2);
      sm.meta[p] = (ghost << 16) | (ghost << 18);
      sm.tile_cnt[p] = ghost;
    }
  }
  __syncthreads();

  // kPre: the rows of the NEXT tile are requested right after the current tile is staged (their
  // registers are free from then on), so the loads fly during the flush and carry steps instead of
  // stalling the next rank step.
  uint32_t key[kScI], val[kScI];
  auto load_tile = [&](int64_t t0) {
    if (t0 + kScTile <= u.row1) {  // full tile: no bounds checks
#pragma unroll
      for (int it = 0; it < kScI; ++it) load_row<kAoS>(in, t0 + it * kScT + tid, key[it], val[it]);
    } else {
#pragma unroll
      for (int it = 0; it < kScI; ++it) {
        const int64_t row = t0 + it * kScT + tid;
        key[it] = 0;
        val[it] = 0;
        if (row < u.row1) load_row<kAoS>(in, row, key[it], val[it]);
      }
    }
  };
  if (kPre && u.row0 < u.row1) load_tile(u.row0);

  for (int64_t t0 = u.row0; t0 < u.row1; t0 += kScTile) {
    // ---- (load,) hash once, rank inside the bucket (ranks continue after the carried rows) ----
    if (!kPre) load_tile(t0);
    uint32_t packed[kScI];  // bucket | rank << 16
    if (t0 + kScTile <= u.row1) {
      if (sel.mask == 0) {
#pragma unroll
        for (int it = 0; it < kScI; ++it) {
          const uint32_t b = part_bucket(wang_hash_u32(key[it]), g.shl, g.bits);
          packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
        }
      } else {
#pragma unroll
        for (int it = 0; it < kScI; ++it) {
          const uint32_t b = bucket_or_skip<false>(key[it], val[it], g, sel);
          packed[it] = b;
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < kScI; ++it) {
        const int64_t row = t0 + it * kScT + tid;
        packed[it] = 0xffffffffu;
        if (row < u.row1) {
          const uint32_t b = bucket_or_skip<false>(key[it], val[it], g, sel);
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
        }
      }
    }
    __syncthreads();

    // ---- one scan for the staged position and the first flushed sector of every bucket ----
    {
      uint32_t mine = 0;
#pragma unroll
      for (int q = 0; q < kScBpt; ++q) {
        const int p = kScBpt * tid + q;
        if (p < P) {
          const uint32_t tot = sm.tile_cnt[p];  // carried + new
          // rows <= 8192 and sectors <= 2816 never carry into each other
          mine += (tot - ((sm.meta[p] >> 16) & 3u)) | ((tot >> 2) << 16);
        }
      }
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) sm.warp_tot[warp] = incl;
      __syncthreads();
      const uint32_t w = lane < kW ? sm.warp_tot[lane] : 0;
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      const uint32_t all = __shfl_sync(0xffffffffu, wi, kW - 1);
      if (tid == 0) sm.n_sec = all >> 16;
      uint32_t excl = __shfl_sync(0xffffffffu, wi - w, warp) + incl - mine;
#pragma unroll
      for (int q = 0; q < kScBpt; ++q) {
        const int p = kScBpt * tid + q;
        if (p < P) {
          // re-read rather than keep live across the scan (the row registers fill the budget)
          const uint32_t tot = sm.tile_cnt[p];
          const uint32_t cg = sm.meta[p] >> 16;  // carried | ghost << 2
          const uint32_t tc = tot - (cg & 3u), ns = tot >> 2;
          const uint32_t start = excl & 0xffffu, so = excl >> 16;
          const int32_t sb = (int32_t)start - (int32_t)(cg & 3u);
          sm.sbase[p] = sb;
          ScDesc d;
          d.sbase = sb - (int32_t)(so * kScRows);
          d.dsec = sm.bd[p].dsec - so;
          sm.bd[p] = d;
          sm.meta[p] = so | (cg << 16);
          if (ns) sm.sec_desc[so] = (uint16_t)(p | (cg << 10));  // the first sector may hold carried / ghost rows
          for (uint32_t i = 1; i < ns; ++i) sm.sec_desc[so + i] = (uint16_t)p;
          excl += tc | (ns << 16);
        }
      }
    }
    __syncthreads();

    // ---- stage the tile sorted by bucket ----
#pragma unroll
    for (int it = 0; it < kScI; ++it) {
      if (packed[it] != 0xffffffffu) {
        const uint32_t b = packed[it] & 0xffffu;
        sm.stage[sm.sbase[b] + (int32_t)(packed[it] >> 16)] = make_uint2(key[it], val[it]);
      }
    }
    __syncthreads();

    if (kPre && t0 + kScTile < u.row1) load_tile(t0 + kScTile);

    // ---- flush whole sectors: four adjacent lanes per 32-byte sector of the destination ----
    {
      const uint32_t n_slot = sm.n_sec * kScRows;
      const uint32_t l = tid & 3;
      constexpr int kUn = 4;  // independent slots in flight per thread: the table -> descriptor -> row
                              // chain is three dependent shared-memory reads
      for (uint32_t t0s = tid; t0s < n_slot; t0s += kUn * kScT) {  // t = 4 * sector + lane in the sector
        uint32_t sd[kUn];
        ScDesc d[kUn];
        uint2 kv[kUn];
#pragma unroll
        for (int k = 0; k < kUn; ++k) {
          const uint32_t t = t0s + k * kScT;
          sd[k] = t < n_slot ? sm.sec_desc[t >> 2] : 0u;  // past the end: bucket 0, never stored
        }
#pragma unroll
        for (int k = 0; k < kUn; ++k) d[k] = sm.bd[sd[k] & 1023u];
#pragma unroll
        for (int k = 0; k < kUn; ++k) {
          const uint32_t t = t0s + k * kScT;
          const uint32_t b = sd[k] & 1023u, c0 = (sd[k] >> 10) & 3u;
          const uint32_t si = t < n_slot ? (uint32_t)(d[k].sbase + (int32_t)t) : 0u;
          const uint2* src = l < c0 ? &sm.carry[b * 3 + l] : &sm.stage[si];
          kv[k] = *src;
        }
#pragma unroll
        for (int k = 0; k < kUn; ++k) {
          const uint32_t t = t0s + k * kScT;
          const uint64_t row = (uint64_t)(d[k].dsec + (t >> 2)) * kScRows + l;
          if (t < n_slot && l >= (sd[k] >> 12)) {  // ghost rows belong to the run before ours
            if (row < cap) st_stream_v2(out + row, kv[k]);
            else if (overflow) *overflow = 1u;
          }
        }
      }
    }
    __syncthreads();

    // ---- hold back the rows that did not complete a sector; advance; reset the counters ----
#pragma unroll
    for (int q = 0; q < kScBpt; ++q) {
      const int p = kScBpt * tid + q;
      if (p < P) {
        const uint32_t tot = sm.tile_cnt[p];
        const uint32_t m = sm.meta[p];
        const uint32_t so = m & 0xffffu, c0 = (m >> 16) & 3u, ghost = m >> 18;
        const uint32_t nsec = tot >> 2, rem = tot & 3u;
        // nothing flushed: append the new rows to the carry; otherwise the tail of the staged rows
        // becomes the carry (carried < 4 <= 4 * nsec)
        const uint32_t first = nsec ? 0u : c0;
        const int32_t from = sm.sbase[p] + (int32_t)(nsec * kScRows);
#pragma unroll
        for (uint32_t i = 0; i < 3; ++i)
          if (i >= first && i < rem) sm.carry[p * 3 + i] = sm.stage[from + (int32_t)i];
        sm.bd[p].dsec += so + nsec;
        sm.meta[p] = (rem << 16) | ((nsec ? 0u : ghost) << 18);
        sm.tile_cnt[p] = rem;
      }
    }
    __syncthreads();
  }

  // ---- end of the unit: the last, partial sector of every bucket ----
  for (int i = tid; i < P * kScRows; i += kScT) {
    const int b = i >> 2;
    const uint32_t l = i & 3;
    const uint32_t m = sm.meta[b];
    if (l >= (m >> 18) && l < ((m >> 16) & 3u)) {
      const uint64_t row = (uint64_t)sm.bd[b].dsec * kScRows + l;
      if (row < cap) st_stream_v2(out + row, sm.carry[b * 3 + l]);
      else if (overflow) *overflow = 1u;
    }
  }
}

// part_off[s*P + p] = first output row of (segment s, bucket p); part_off[nseg*P] = total rows.
// ---- whole-sector scatter, second generation: quad-aligned regions ("quads") -----------------------
// The source profile of part_scatter_sectors_kernel (profiles/r2_scatter_probe_phases.md) showed the
// flush as HALF of its instructions: per stored row it looked up the sector's bucket, unpacked the
// bucket's descriptor, chose between the carry array and the stage, rebuilt the stage index and the
// 64-bit destination row. This kernel removes the choices instead of speeding them up:
//   * a bucket's region of the stage starts on a multiple of four slots and holds (rows carried from
//     the previous tile ++ this tile's rows): the carried rows are copied INTO the stage by the
//     bucket's thread during the scan step, so the flush reads stage[t] and nothing else;
//   * every quad of stage slots belongs to one bucket and maps onto one aligned 32-byte sector of the
//     output, so slot t goes to sector desc[bucket].dsec + t / 4, lane t % 4: one add, one wide
//     multiply-add for the address;
//   * "is this slot stored?" is lo <= t < hi with lo / hi precomputed per bucket (the rows of the
//     first sector that belong to the neighbouring run, and the tail that does not fill a sector).
// One 1024-thread CTA per SM, 14 rows per thread per tile (the stage, its padding and the carry
// array fill the 227 KB of shared memory), one bucket per thread in the per-bucket steps.
constexpr int kQdT = 1024, kQdI = 14;
constexpr int kQdTile = kQdT * kQdI;                                  // 14336 rows
constexpr int kQdSlots = kQdTile + 6 * (1 << kPartMaxBits);           // + carried rows + padding to quads
constexpr int kQdUn = 4;                                              // slots in flight per thread in the flush
constexpr int kQdRound = kQdT * kQdUn;                                // slots one flush round covers
static_assert(kQdSlots % kQdRound == 0, "the stage is a whole number of flush rounds");
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {  // explicit shared-window loads: no generic -> shared conversion
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
struct QdDesc {
  uint32_t dsec;     // destination sector of the region's first quad, minus that quad's index
  uint16_t lo, hi;   // slots [lo, hi) of the stage are stored
};
struct QdSmem {
  uint2 stage[kQdSlots];                         // 160 KB
  uint2 carry[(1 << kPartMaxBits) * 3];          // 24 KB: the rows that did not fill a sector
  QdDesc desc[(1 << kPartMaxBits) + 1];          // 8 KB; the last entry (lo == hi) stores nothing: padding quads point at it
  uint32_t tile_cnt[1 << kPartMaxBits];          // carried (incl. ghost) + rows ranked so far
  uint32_t qstart[1 << kPartMaxBits];            // first stage slot of the bucket's region
  uint32_t dsec_next[1 << kPartMaxBits];         // next unwritten sector of the (unit, bucket) run
  uint16_t quad_bucket[kQdSlots / 4];            // 10 KB: bucket of every quad
  uint8_t carried[1 << kPartMaxBits];            // rows in carry[] (ghost rows included)
  uint8_t ghost[1 << kPartMaxBits];              // leading slots of the run's first sector that are not ours
  uint32_t warp_tot[kQdT / 32];
  uint32_t n_slots;
};
static_assert(sizeof(QdSmem) <= 227 * 1024, "QdSmem must fit the opt-in shared memory of one CTA");

template <bool kAoS, bool kPre>
__global__ void __launch_bounds__(kQdT, 1)
part_scatter_quads_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                          const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                          PartGeom g, const uint64_t* __restrict__ scanned, uint2* __restrict__ out,
                          int64_t out_cap, unsigned int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem[];
  QdSmem& sm = *reinterpret_cast<QdSmem*>(smem);
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P);
  if (!u.valid) return;
  constexpr int kW = kQdT / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mine = (int)tid < P;  // this thread's bucket in the per-bucket steps

  if (mine) {
    const uint64_t pos = scanned[u.hbase + (int64_t)tid * u.ustride];  // first output row of (unit, bucket)
    const uint32_t gh = (uint32_t)(pos & (kScRows - 1));               // rows of that sector that are not ours
    sm.dsec_next[tid] = (uint32_t)(pos >> 2);
    sm.ghost[tid] = (uint8_t)gh;
    sm.carried[tid] = (uint8_t)gh;
    sm.tile_cnt[tid] = gh;
  }
  __syncthreads();

  if (tid == 0) {
    QdDesc none;
    none.dsec = 0;
    none.lo = 0;
    none.hi = 0;
    sm.desc[1 << kPartMaxBits] = none;
  }
  // per-thread constants of the flush: lane in the sector, its byte offset, the last sector it may store,
  // and the shared-window addresses of the three tables it reads
  const uint32_t l = tid & 3;
  unsigned char* out_l = reinterpret_cast<unsigned char*>(out) + 8 * l;
  const uint32_t qb_base = (uint32_t)__cvta_generic_to_shared(sm.quad_bucket);
  const uint32_t desc_base = (uint32_t)__cvta_generic_to_shared(sm.desc);
  const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(sm.stage);
  const uint64_t cap_rows = (uint64_t)out_cap;
  const uint32_t cap_sec = cap_rows > l ? (uint32_t)min((cap_rows - l + 3) >> 2, (uint64_t)0xffffffffu) : 0u;

  uint32_t key[kQdI], val[kQdI];
  auto load_tile = [&](int64_t t0) {
    if (t0 + kQdTile <= u.row1) {  // full tile: no bounds checks
#pragma unroll
      for (int it = 0; it < kQdI; ++it) load_row<kAoS>(in, t0 + it * kQdT + tid, key[it], val[it]);
    } else {
#pragma unroll
      for (int it = 0; it < kQdI; ++it) {
        const int64_t row = t0 + it * kQdT + tid;
        key[it] = 0;
        val[it] = 0;
        if (row < u.row1) load_row<kAoS>(in, row, key[it], val[it]);
      }
    }
  };
  if (kPre && u.row0 < u.row1) load_tile(u.row0);

  for (int64_t t0 = u.row0; t0 < u.row1; t0 += kQdTile) {
    // ---- (load,) hash once, rank inside the bucket (ranks continue after the carried rows) ----
    if (!kPre) load_tile(t0);
    uint32_t packed[kQdI];  // bucket | rank << 16
    if (t0 + kQdTile <= u.row1 && sel.mask == 0) {
#pragma unroll
      for (int it = 0; it < kQdI; ++it) {
        const uint32_t b = part_bucket(wang_hash_u32(key[it]), g.shl, g.bits);
        packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
      }
    } else {
#pragma unroll
      for (int it = 0; it < kQdI; ++it) {
        const int64_t row = t0 + it * kQdT + tid;
        packed[it] = 0xffffffffu;
        if (row < u.row1) {
          const uint32_t b = bucket_or_skip<false>(key[it], val[it], g, sel);
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&sm.tile_cnt[b], 1u) << 16);
        }
      }
    }
    __syncthreads();

    // ---- per bucket: region (rounded up to whole quads), descriptor, carried rows into the stage ----
    uint32_t tot = 0, c0 = 0, gh = 0;
    if (mine) {
      tot = sm.tile_cnt[tid];
      c0 = sm.carried[tid];
      gh = sm.ghost[tid];
    }
    const uint32_t region = (tot + 3u) & ~3u;
    uint32_t incl = region;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    {
      const uint32_t w = sm.warp_tot[lane];  // kW == 32
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      const uint32_t all = __shfl_sync(0xffffffffu, wi, kW - 1);
      if (tid == 0) sm.n_slots = all;
      const uint32_t q0 = __shfl_sync(0xffffffffu, wi - w, warp) + incl - region;  // first slot of the region
      if (mine) {
        const uint32_t nsec = tot >> 2;
        sm.qstart[tid] = q0;
        QdDesc d;
        d.dsec = sm.dsec_next[tid] - (q0 >> 2);
        d.lo = (uint16_t)(q0 + gh);
        d.hi = (uint16_t)(q0 + nsec * kScRows);
        sm.desc[tid] = d;
        // rows beyond the output's capacity are dropped by the flush (its store is predicated on the
        // sector); the flag is raised here, once per bucket and tile, not per row
        if (overflow && ((uint64_t)sm.dsec_next[tid] + nsec) * kScRows > cap_rows) *overflow = 1u;
        for (uint32_t e = gh; e < c0; ++e) sm.stage[q0 + e] = sm.carry[tid * 3 + e];
        for (uint32_t q = 0; q < (region >> 2); ++q) sm.quad_bucket[(q0 >> 2) + q] = (uint16_t)tid;
      }
      // quads between the last region and the end of the last flush round belong to the sentinel
      for (uint32_t q = (all >> 2) + tid; q < (((all + kQdRound - 1) / kQdRound * kQdRound) >> 2); q += kQdT)
        sm.quad_bucket[q] = (uint16_t)(1 << kPartMaxBits);
    }
    __syncthreads();

    // ---- stage the tile's rows behind the carried ones ----
#pragma unroll
    for (int it = 0; it < kQdI; ++it) {
      if (packed[it] != 0xffffffffu)
        sm.stage[sm.qstart[packed[it] & 0xffffu] + (packed[it] >> 16)] = make_uint2(key[it], val[it]);
    }
    __syncthreads();

    if (kPre && t0 + kQdTile < u.row1) load_tile(t0 + kQdTile);

    // ---- flush: slot t -> sector desc.dsec + t / 4, lane t % 4; four adjacent lanes = one sector ----
    // Written with explicit shared-window addresses and predicated stores: the compiler's version of
    // this loop spent 35 instructions per slot on re-deriving the shared base, the output base and on
    // four divergent branches per round (profiles/r2_scatter_probe_phases.md); this one spends ~14.
    {
      const uint32_t rounds = (sm.n_slots + kQdRound - 1) / kQdRound;
      uint32_t t = tid;
      uint64_t out_l64 = reinterpret_cast<uint64_t>(out_l);
      asm volatile("" : "+l"(out_l64));  // opaque: keep the base in a register pair instead of re-deriving it per store
      for (uint32_t r = 0; r < rounds; ++r, t += kQdRound) {
        uint32_t b[kQdUn];
        uint2 dd[kQdUn], kv[kQdUn];
#pragma unroll
        for (int k = 0; k < kQdUn; ++k) b[k] = lds_u16(qb_base + (((t + k * kQdT) >> 2) << 1));
#pragma unroll
        for (int k = 0; k < kQdUn; ++k) {
          dd[k] = lds_v2(desc_base + b[k] * 8u);
          kv[k] = lds_v2(stage_base + (t + k * kQdT) * 8u);
        }
#pragma unroll
        for (int k = 0; k < kQdUn; ++k) {
          const uint32_t tk = t + k * kQdT;
          const uint32_t sector = dd[k].x + (tk >> 2);
          const bool ok = tk >= (dd[k].y & 0xffffu) && tk < (dd[k].y >> 16) && sector < cap_sec;
          if (ok) st_stream_v2(reinterpret_cast<uint2*>(out_l64 + (uint64_t)sector * 32u), kv[k]);
        }
      }
    }
    __syncthreads();

    // ---- per bucket: the tail that did not fill a sector is carried; advance; reset the counter ----
    if (mine) {
      const uint32_t nsec = tot >> 2, rem = tot & 3u;
      const uint32_t from = sm.qstart[tid] + nsec * kScRows;
      // nothing flushed: slots below ghost are still not ours; otherwise the tail starts a new sector
      const uint32_t first = nsec ? 0u : gh;
      for (uint32_t i = first; i < rem; ++i) sm.carry[tid * 3 + i] = sm.stage[from + i];
      sm.dsec_next[tid] += nsec;
      sm.carried[tid] = (uint8_t)rem;
      if (nsec) sm.ghost[tid] = 0;
      sm.tile_cnt[tid] = rem;
    }
    __syncthreads();
  }

  // ---- end of the unit: the last, partial sector of every bucket ----
  for (int i = tid; i < P * kScRows; i += kQdT) {
    const int b = i >> 2;
    const uint32_t li = i & 3;
    if (li >= sm.ghost[b] && li < sm.carried[b]) {
      const uint64_t row = (uint64_t)sm.dsec_next[b] * kScRows + li;
      if (row < cap_rows) st_stream_v2(out + row, sm.carry[b * 3 + li]);
      else if (overflow) *overflow = 1u;
    }
  }
}


// ---- whole-sector scatter, third generation: the copy engine flushes ("bulk") ----------------------
// The sector and quad kernels above spend most of their instructions in the flush: every stored row
// costs a bucket lookup, a descriptor read, an address computation and a predicated 8-byte store
// (profiles/r2_launches_join_sf1024.csv: 87 lane-instructions per row, 45 % issue utilisation, 3.2 TB/s
// where the memory system takes the same run lengths at 5.2 TB/s, tools/microbench_scatter_ceiling.cu).
// Here the flush is ONE instruction per (bucket, tile): the bucket's region of the stage — quad-aligned,
// carried rows first, as in the quads kernel — is handed to the copy engine as one shared -> global bulk
// copy of its whole sectors (cp.async.bulk, UBLKCP in SASS), 16-byte aligned at both ends by
// construction. No sector table, no descriptors, no per-row stores. What is left per row is the load,
// the hash, one shared-memory atomic for the rank and one shared-memory store into the stage.
//   * thread p owns bucket p in the per-bucket steps and keeps the bucket's state across tiles in registers
//     (next output sector, number of held-back rows, how many of the first sector's slots belong to the
//     neighbouring run: "ghost") and in a private column of a small carry array (the at most 3 rows that
//     did not complete a sector);
//   * rank counters are double-buffered: the counter of the next tile is initialised (to the number of
//     carried rows) in the per-bucket step of this tile, which removes the barrier at the end of a tile:
//     four barriers per tile (rank | scan | regions | staged);
//   * the copy engine reads the stage asynchronously: cp.async.bulk.wait_group.read before the barrier
//     that precedes the next tile's staging, a full tile of work after the copies were issued;
//   * a (unit, bucket) run's first sector, when shared with the neighbouring run, and its last, partial
//     sector are written with plain 8-byte stores by the bucket's thread.
// Two shapes: kBkT = 1024 threads, one CTA per SM, 16384-row tiles, up to 1024 buckets; kBkT = 512 threads,
// two CTAs per SM, 8192-row tiles, up to 512 buckets — the same 16 rows per bucket and tile at a fan-out
// of 512, with one CTA's rank / scan / stage steps running under the other's loads and copies.
template <int kBkT>
struct BkSmemT {
  static constexpr int kBuckets = kBkT;                  // one bucket per thread in the per-bucket steps
  static constexpr int kTile = kBkT * 16;
  static constexpr int kSlots = kTile + 6 * kBuckets;    // + carried rows + padding to quads
  uint2 stage[kSlots];                           // 176 KB / 88 KB
  uint32_t cnt[2][kBuckets];                     // rank counters of this tile / the next tile
  uint32_t qstart[kBuckets];                     // first stage slot of the bucket's region
  uint2 carry[3][kBuckets];                      // the rows that did not complete a sector (private to the bucket's thread)
  uint32_t warp_tot[kBkT / 32];
};
static_assert(sizeof(BkSmemT<1024>) <= 227 * 1024, "must fit the opt-in shared memory of one CTA");
static_assert(2 * (sizeof(BkSmemT<512>) + 1024) <= 227 * 1024, "two CTAs of the small shape per SM");

__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// kPeer: the fused multi-GPU shuffle. Bucket p of THIS rank starts at bucket_addr[p], a byte address inside
// another GPU's receive buffer (the bulk copies then cross NVLink), this unit's run where the units before
// it in the segment end; sectors are counted from the address itself; the receive capacity was checked
// by the plan kernel (abort_flag); a grid smaller than the number of work units walks them (CTA budget).
template <bool kAoS, bool kPre, int kBkT, bool kPeer = false>
__global__ void __launch_bounds__(kBkT, (kBkT >= 1024 ? 1 : 2))
part_scatter_bulk_kernel(PartInput in, const int64_t* __restrict__ seg_off,
                         const int64_t* __restrict__ unit_first, int64_t nseg, int64_t unit_rows,
                         PartGeom g, const uint64_t* __restrict__ scanned, uint2* __restrict__ out,
                         int64_t out_cap, unsigned int* __restrict__ overflow,
                         const uint64_t* __restrict__ bucket_addr = nullptr,
                         const int64_t* __restrict__ abort_flag = nullptr) {
  extern __shared__ __align__(16) unsigned char smem[];  // the window itself starts 1024-byte aligned
  using BkSmem = BkSmemT<kBkT>;
  constexpr int kBkI = 16, kBkTile = BkSmem::kTile;
  BkSmem& sm = *reinterpret_cast<BkSmem*>(smem);
  if (kPeer && abort_flag && *abort_flag) return;  // the exchange was called off on the device: store nothing
  const int P = 1 << g.bits;
  const SliceSel sel = slice_sel(g);
  constexpr int kW = kBkT / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mine = (int)tid < P;  // this thread's bucket in the per-bucket steps
  const uint64_t cap_rows = kPeer ? ~0ull : (uint64_t)out_cap;
  const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(sm.stage);
  for (int64_t unit_id = blockIdx.x;; unit_id += gridDim.x) {
  const Unit u = find_unit(seg_off, unit_first, nseg, unit_rows, P, unit_id);
  if (!u.valid) return;

  // per-bucket state, in registers of the bucket's thread
  uint32_t dsec_next = 0;  // next unwritten sector of the (unit, bucket) run, counted from `out` (kPeer: from obase)
  unsigned char* obase = reinterpret_cast<unsigned char*>(out);  // kPeer: the bucket's run start, sector-aligned down
  uint32_t gh = 0;         // leading slots of the run's first sector that belong to the neighbouring run
  uint32_t c0 = 0;         // rows held back (ghost slots included)
  if (mine) {
    const uint64_t pos = scanned[u.hbase + (int64_t)tid * u.ustride];  // first output row of (unit, bucket)
    if (kPeer) {
      const uint64_t a0 = bucket_addr[tid] + 8ull * (pos - scanned[u.hbase0 + (int64_t)tid * u.ustride]);
      gh = (uint32_t)((a0 >> 3) & (kScRows - 1));
      obase = reinterpret_cast<unsigned char*>(a0 - 8ull * gh);
      dsec_next = 0;
    } else {
      gh = (uint32_t)(pos & (kScRows - 1));
      dsec_next = (uint32_t)(pos >> 2);
    }
    c0 = gh;
    sm.cnt[0][tid] = gh;
  }
  __syncthreads();

  uint32_t key[kBkI], val[kBkI];
  auto load_tile = [&](int64_t t0) {
#ifdef B2_LAB
    if (g.lab & 1) {  // experiment: what the pass costs without its DRAM reads (the output is wrong on purpose)
#pragma unroll
      for (int it = 0; it < kBkI; ++it) {
        val[it] = (uint32_t)(t0 + it * kBkT + tid);
        key[it] = val[it] * 2654435761u;
      }
      return;
    }
#endif
    if (t0 + kBkTile <= u.row1) {  // full tile: no bounds checks
#pragma unroll
      for (int it = 0; it < kBkI; ++it) load_row<kAoS>(in, t0 + it * kBkT + tid, key[it], val[it]);
    } else {
#pragma unroll
      for (int it = 0; it < kBkI; ++it) {
        const int64_t row = t0 + it * kBkT + tid;
        key[it] = 0;
        val[it] = 0;
        if (row < u.row1) load_row<kAoS>(in, row, key[it], val[it]);
      }
    }
  };
  if (kPre && u.row0 < u.row1) load_tile(u.row0);

  uint32_t par = 0;
  for (int64_t t0 = u.row0; t0 < u.row1; t0 += kBkTile, par ^= 1u) {
    // The tile after this one is requested into L2 now (one instruction, no registers): the register
    // loads that follow this tile's staging then hit L2 instead of waiting for DRAM, whose reads run
    // under the rank / scan / stage steps (tools/part_lab.py with B2_LAB_SCATTER: the exposed loads were
    // a quarter of the pass).
    if (kPre && tid == 0 && t0 + kBkTile < u.row1) {
      const int64_t n0 = t0 + kBkTile, nn = min((int64_t)kBkTile, u.row1 - n0);
      if (kAoS) {
        l2_prefetch(in.pairs + n0, nn * 8);
      } else {
        l2_prefetch(in.keys + n0, nn * 4);
        if (in.vals) l2_prefetch(in.vals + n0, nn * 4);
      }
    }
    // ---- (load,) hash once, rank inside the bucket (ranks continue after the held-back rows) ----
    if (!kPre) load_tile(t0);
    uint32_t* cnt = sm.cnt[par];
    uint32_t packed[kBkI];  // bucket | rank << 16
    if (t0 + kBkTile <= u.row1 && sel.mask == 0) {
#pragma unroll
      for (int it = 0; it < kBkI; ++it) {
        const uint32_t b = part_bucket(wang_hash_u32(key[it]), g.shl, g.bits);
        packed[it] = b | (atomicAdd(&cnt[b], 1u) << 16);
      }
    } else {
#pragma unroll
      for (int it = 0; it < kBkI; ++it) {
        const int64_t row = t0 + it * kBkT + tid;
        packed[it] = 0xffffffffu;
        if (row < u.row1) {
          const uint32_t b = bucket_or_skip<false>(key[it], val[it], g, sel);
          if (b != 0xffffffffu) packed[it] = b | (atomicAdd(&cnt[b], 1u) << 16);
        }
      }
    }
    __syncthreads();

    // ---- per bucket: region of the stage (rounded up to whole quads), next tile's counter ----
    const uint32_t tot = mine ? cnt[tid] : 0u;  // held back (incl. ghost) + this tile's rows
    const uint32_t region = (tot + 3u) & ~3u;
    uint32_t incl = region;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    if (mine) sm.cnt[par ^ 1u][tid] = tot & 3u;  // what this tile will hold back
    bulk_wait_read();  // the copies of the previous tile have read the stage (issued a tile of work ago)
    __syncthreads();
    uint32_t q0;
    {
      const uint32_t w = lane < kW ? sm.warp_tot[lane] : 0u;
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      q0 = __shfl_sync(0xffffffffu, wi - w, warp) + incl - region;  // first slot of the region
      if (mine) sm.qstart[tid] = q0;
    }
    __syncthreads();

    // ---- stage the tile's rows behind the held-back ones ----
#pragma unroll
    for (int it = 0; it < kBkI; ++it) {
      if (packed[it] != 0xffffffffu)
        sm.stage[sm.qstart[packed[it] & 0xffffu] + (packed[it] >> 16)] = make_uint2(key[it], val[it]);
    }
    if (mine) {
      for (uint32_t e = gh; e < c0; ++e) sm.stage[q0 + e] = sm.carry[e][tid];
    }
    fence_proxy_async();  // the copy engine reads what this thread staged
    __syncthreads();

    if (kPre && t0 + kBkTile < u.row1) load_tile(t0 + kBkTile);

    // ---- flush: one bulk copy per bucket; hold back the tail ----
    if (mine) {
      const uint32_t nsec = tot >> 2, rem = tot & 3u;
      if (nsec) {
        if (((uint64_t)dsec_next + nsec) * kScRows > cap_rows) {
          if (overflow) *overflow = 1u;  // rows beyond the output's capacity are dropped
        } else {
          uint32_t first = 0;
          if (gh) {  // the run's first sector is shared with the neighbouring run: plain stores
            for (uint32_t e = gh; e < kScRows; ++e)
              st_stream_v2(reinterpret_cast<uint2*>(obase) + ((uint64_t)dsec_next * kScRows + e), sm.stage[q0 + e]);
            first = 1;
          }
#ifdef B2_LAB
          if (g.lab & 2) first = nsec;  // experiment: what the pass costs without its DRAM writes
#endif
          if (nsec > first) {
            bulk_store(obase + ((uint64_t)dsec_next + first) * 32u,
                       stage_base + (q0 + first * kScRows) * 8u, (nsec - first) * 32u);
            bulk_commit();
          }
        }
        dsec_next += nsec;
        gh = 0;
      }
      const uint32_t from = q0 + nsec * kScRows;
      for (uint32_t e = 0; e < rem; ++e) sm.carry[e][tid] = sm.stage[from + e];
      c0 = rem;
    }
  }

  // ---- end of the unit: the last, partial sector of every bucket ----
  if (mine && c0 > gh) {
    const uint64_t row = (uint64_t)dsec_next * kScRows;
    if (row + c0 > cap_rows) {
      if (overflow) *overflow = 1u;
    } else {
      for (uint32_t e = gh; e < c0; ++e) st_stream_v2(reinterpret_cast<uint2*>(obase) + row + e, sm.carry[e][tid]);
    }
  }
  bulk_wait_read();  // shared memory must outlive the copy engine's reads
  if (!kPeer) return;  // one work unit per CTA
  __syncthreads();     // the next unit re-initialises the counters and the stage
  }  // units of this CTA
}

// Boundaries are clamped to `clamp` (the capacity of the pass's output): when a skewed hash-space
// slice overflows its buffer the scatter drops the rows beyond it and raises *overflow; whoever reads
// the output by these boundaries (a second pass, the probe) must not run past the buffer either.
__global__ void part_offsets_kernel(const uint64_t* __restrict__ scanned,
                                    const int64_t* __restrict__ unit_first, int64_t nseg, int P,
                                    int64_t clamp, int64_t* __restrict__ part_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = nseg * P;
  if (i > n) return;
  if (i == n) {
    part_off[n] = min((int64_t)scanned[unit_first[nseg] * P], clamp);
    return;
  }
  const int64_t s = i / P, p = i - s * P;
  const int64_t ustride = unit_first[s + 1] - unit_first[s];
  part_off[i] = min((int64_t)scanned[unit_first[s] * P + p * ustride], clamp);
}

size_t scatter_smem_bytes(int bits, int tile_rows) {
  const size_t P = (size_t)1 << bits;
  const size_t tile = (size_t)tile_rows;
  return sizeof(uint2) * tile + P * 8 + P * 8 + P * 4 + P * 4 + sizeof(uint16_t) * tile;
}

int64_t choose_unit_rows(int64_t n, int64_t nseg) {
  int64_t tiles = n / ((int64_t)kPartTile * 2048);
  const int64_t avg_seg_tiles = nseg > 0 ? (n / nseg) / kPartTile : 1;
  if (tiles > avg_seg_tiles) tiles = avg_seg_tiles;
  if (tiles < 1) tiles = 1;
  if (tiles > 32) tiles = 32;
  return tiles * kPartTile;
}

struct PassLayout {
  int64_t unit_rows, max_units, n_entries;  // n_entries includes the trailing total slot
  size_t off_unit_first, off_hist, off_scanned, off_scanws, total;
};

PassLayout pass_layout(int64_t n, int64_t nseg, int bits) {
  PassLayout L;
  L.unit_rows = choose_unit_rows(n, nseg);
  L.max_units = n / L.unit_rows + nseg;
  L.n_entries = L.max_units * ((int64_t)1 << bits) + 1;
  size_t o = 0;
  L.off_unit_first = o; o += b2_align_up((size_t)(nseg + 1) * 8, 256);
  L.off_hist = o;       o += b2_align_up((size_t)L.n_entries * 4, 256);
  L.off_scanned = o;    o += b2_align_up((size_t)L.n_entries * 8, 256);
  L.off_scanws = o;     o += b2_align_up(b2_scan_ws_bytes(L.n_entries), 256);
  L.total = o;
  return L;
}

// Kernel-selection knobs live in the ctx (b2_ctx_set_tunable), not in process-global state:
//   B2_TUNE_SCATTER_SECTORS_MIN_BITS  fan-out (log2) from which the whole-sector scatter kernel is used (default 9,
//                                     measured: tools/join_lab.py smem:0:b)
//   B2_TUNE_SCATTER_PREFETCH          request the next tile's rows while the current one is flushed (default on)
//   B2_TUNE_SCATTER_SHAPE             threads x rows per thread x CTAs per SM of the plain scatter kernel

template <bool kAoS, int kT, int kI, int kCtas, bool kValPred = false, bool kPre = false>
int launch_scatter(b2_ctx* ctx, int64_t units, int bits, cudaStream_t s, const PartInput& in,
                   const int64_t* d_seg_off, const int64_t* unit_first, int64_t nseg, int64_t unit_rows,
                   const PartGeom& g, const uint64_t* scanned, uint2* d_out, int64_t out_cap,
                   unsigned int* d_overflow) {
  static const int seen = b2_new_site();
  if (b2_first_use_on_device(ctx, seen)) {
    B2_CUDA_OK(ctx, cudaFuncSetAttribute(part_scatter_kernel<kAoS, kT, kI, kCtas, kValPred, kPre>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)scatter_smem_bytes(kPartMaxBits, kT * kI)));
  }
  part_scatter_kernel<kAoS, kT, kI, kCtas, kValPred, kPre><<<(unsigned)units, kT, scatter_smem_bytes(bits, kT * kI), s>>>(
      in, d_seg_off, unit_first, nseg, unit_rows, g, scanned, d_out, out_cap, nullptr, d_overflow);
  return B2_OK;
}

// Phase 1 of a pass: unit table, histogram, flat scan, partition boundaries. Leaves the scanned
// histogram in the workspace for part_scatter_impl (same n / nseg / geometry / workspace).
template <bool kAoS>
int part_count_impl(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off,
                    int64_t nseg, const PartGeom& g, int64_t* d_part_off, void* d_ws, size_t ws_bytes,
                    cudaStream_t s, int64_t clamp) {
  const int P = 1 << g.bits;
  const PassLayout L = pass_layout(n, nseg, g.bits);
  if (ws_bytes < L.total) return b2_set_error(ctx, B2_ERR_WORKSPACE, "partition pass", "workspace");
  char* base = static_cast<char*>(d_ws);
  int64_t* unit_first = reinterpret_cast<int64_t*>(base + L.off_unit_first);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + L.off_hist);
  uint64_t* scanned = reinterpret_cast<uint64_t*>(base + L.off_scanned);
  B2_REQUIRE(ctx, L.max_units < (1ll << 31), "too many work units");

  part_unit_table_kernel<<<1, 1024, 0, s>>>(d_seg_off, nseg, L.unit_rows, unit_first);
  B2_LAUNCH_CHECK(ctx, "part_unit_table_kernel");
  B2_CUDA_OK(ctx, cudaMemsetAsync(hist, 0, (size_t)L.n_entries * 4, s));
  if (L.max_units > 0) {
    if (g.val_pred)
      part_hist_kernel<kAoS, true><<<(unsigned)L.max_units, kThreads, 0, s>>>(
          in, d_seg_off, unit_first, nseg, L.unit_rows, g, hist);
    else
      part_hist_kernel<kAoS, false><<<(unsigned)L.max_units, kThreads, 0, s>>>(
          in, d_seg_off, unit_first, nseg, L.unit_rows, g, hist);
    B2_LAUNCH_CHECK(ctx, "part_hist_kernel");
  }
  B2_RETURN_NOT_OK(b2_exclusive_scan_u32_u64(ctx, hist, scanned, L.n_entries, base + L.off_scanws,
                                             ws_bytes - L.off_scanws, s));
  if (d_part_off) {
    const int64_t cnt = nseg * P + 1;
    part_offsets_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(scanned, unit_first, nseg, P, clamp,
                                                                      d_part_off);
    B2_LAUNCH_CHECK(ctx, "part_offsets_kernel");
  }
  return B2_OK;
}

// Phase 2: scatter, to d_out (local) or to the per-bucket byte addresses in d_bucket_addr (peer).
template <bool kAoS>
int part_scatter_impl(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off,
                      int64_t nseg, const PartGeom& g_in, uint2* d_out, int64_t out_cap,
                      const uint64_t* d_bucket_addr, unsigned int* d_overflow, void* d_ws,
                      size_t ws_bytes, cudaStream_t s, const int64_t* d_abort) {
  const PassLayout L = pass_layout(n, nseg, g_in.bits);
  if (ws_bytes < L.total) return b2_set_error(ctx, B2_ERR_WORKSPACE, "partition pass", "workspace");
  char* base = static_cast<char*>(d_ws);
  const int64_t* unit_first = reinterpret_cast<const int64_t*>(base + L.off_unit_first);
  const uint64_t* scanned = reinterpret_cast<const uint64_t*>(base + L.off_scanned);
#ifdef B2_LAB
  PartGeom g = g_in;
  if (const char* e = getenv("B2_LAB_SCATTER")) g.lab = atoi(e);
#else
  const PartGeom& g = g_in;
#endif
  if (L.max_units > 0) {
    // whole-sector scatter: local destinations, 32-byte aligned output, fan-out at or above the threshold
    const bool sectors = !d_bucket_addr && !g.val_pred && (reinterpret_cast<uintptr_t>(d_out) & 31) == 0 &&
                         (uint64_t)out_cap < (1ull << 34) && g.bits >= ctx->tune[B2_TUNE_SCATTER_SECTORS_MIN_BITS];
    if (sectors) {
      const int shape = ctx->tune[B2_TUNE_SCATTER_SECTOR_TILE];
      const bool big = shape == 1;
      const bool pre = ctx->tune[B2_TUNE_SCATTER_PREFETCH] != 0;
      // the instantiations share one function-pointer type, so each gets its own site id
      static const int sites[10] = {b2_new_site(), b2_new_site(), b2_new_site(), b2_new_site(), b2_new_site(),
                                    b2_new_site(), b2_new_site(), b2_new_site(), b2_new_site(), b2_new_site()};
      auto go = [&](auto kernel, int site, int threads, size_t smem_bytes) -> int {
        if (b2_first_use_on_device(ctx, site))
          B2_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        kernel<<<(unsigned)L.max_units, threads, smem_bytes, s>>>(in, d_seg_off, unit_first, nseg, L.unit_rows, g,
                                                                 scanned, d_out, out_cap, d_overflow);
        return B2_OK;
      };
      if (shape == 3) {  // quad-aligned regions flushed by the copy engine: one bulk copy per (bucket, tile)
        auto go_bulk = [&](auto kernel, int site, int threads, size_t smem_bytes) -> int {
          if (b2_first_use_on_device(ctx, site))
            B2_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
          kernel<<<(unsigned)L.max_units, threads, smem_bytes, s>>>(in, d_seg_off, unit_first, nseg, L.unit_rows, g,
                                                                   scanned, d_out, out_cap, d_overflow, nullptr, nullptr);
          return B2_OK;
        };
        const bool small = g.bits <= 9 && ctx->tune[B2_TUNE_SCATTER_SHAPE] != 8;  // two 512-thread CTAs per SM
        if (small) {
          if (pre) B2_RETURN_NOT_OK(go_bulk(part_scatter_bulk_kernel<kAoS, true, 512>, sites[8], 512, sizeof(BkSmemT<512>)));
          else B2_RETURN_NOT_OK(go_bulk(part_scatter_bulk_kernel<kAoS, false, 512>, sites[9], 512, sizeof(BkSmemT<512>)));
        } else {
          if (pre) B2_RETURN_NOT_OK(go_bulk(part_scatter_bulk_kernel<kAoS, true, 1024>, sites[6], 1024, sizeof(BkSmemT<1024>)));
          else B2_RETURN_NOT_OK(go_bulk(part_scatter_bulk_kernel<kAoS, false, 1024>, sites[7], 1024, sizeof(BkSmemT<1024>)));
        }
      } else if (shape == 2) {  // quad-aligned regions: the carried rows live in the stage, the flush is three reads and a store
        if (pre) B2_RETURN_NOT_OK(go(part_scatter_quads_kernel<kAoS, true>, sites[4], kQdT, sizeof(QdSmem)));
        else B2_RETURN_NOT_OK(go(part_scatter_quads_kernel<kAoS, false>, sites[5], kQdT, sizeof(QdSmem)));
      } else if (big) {
        if (pre) B2_RETURN_NOT_OK(go(part_scatter_sectors_kernel<kAoS, true, 1024, 16>, sites[0], 1024, sizeof(ScSmemT<1024, 16>)));
        else B2_RETURN_NOT_OK(go(part_scatter_sectors_kernel<kAoS, false, 1024, 16>, sites[1], 1024, sizeof(ScSmemT<1024, 16>)));
      } else {
        if (pre) B2_RETURN_NOT_OK(go(part_scatter_sectors_kernel<kAoS, true, 512, 16>, sites[2], 512, sizeof(ScSmemT<512, 16>)));
        else B2_RETURN_NOT_OK(go(part_scatter_sectors_kernel<kAoS, false, 512, 16>, sites[3], 512, sizeof(ScSmemT<512, 16>)));
      }
      B2_LAUNCH_CHECK(ctx, "part_scatter_sectors_kernel");
      return B2_OK;
    }
    if (d_bucket_addr && ctx->tune[B2_TUNE_PEER_SCATTER_KERNEL] == 1 && !g.val_pred) {
      // peer destinations through the copy engine: whole 32-byte sectors, one bulk copy per (bucket, tile)
      static const int site_a = b2_new_site(), site_s = b2_new_site();
      const int budget = ctx->tune[B2_TUNE_PEER_SCATTER_CTAS];
      const int64_t grid = budget > 0 ? std::min<int64_t>(L.max_units, budget) : L.max_units;
      auto kernel = part_scatter_bulk_kernel<kAoS, true, 1024, true>;
      if (b2_first_use_on_device(ctx, kAoS ? site_a : site_s))
        B2_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(BkSmemT<1024>)));
      kernel<<<(unsigned)grid, 1024, sizeof(BkSmemT<1024>), s>>>(in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned,
                                                                nullptr, 0, nullptr, d_bucket_addr, d_abort);
      B2_LAUNCH_CHECK(ctx, "part_scatter_bulk_kernel (peer)");
      return B2_OK;
    }
    if (d_bucket_addr) {  // peer destinations: whole 128-byte lines only (line carry)
      static const int seen = b2_new_site();  // one table per template instantiation
      if (b2_first_use_on_device(ctx, seen)) {
        B2_CUDA_OK(ctx, cudaFuncSetAttribute(part_scatter_lines_kernel<kAoS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(LcSmem)));
      }
      const int budget = ctx->tune[B2_TUNE_PEER_SCATTER_CTAS];
      const int64_t grid = budget > 0 ? std::min<int64_t>(L.max_units, budget) : L.max_units;
      part_scatter_lines_kernel<kAoS><<<(unsigned)grid, kLcThreads, sizeof(LcSmem), s>>>(
          in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_bucket_addr, d_abort);
    } else if (g.val_pred) {  // pushed-down value predicate: one shape, the default one
      if (ctx->tune[B2_TUNE_SCATTER_PREFETCH])
        B2_RETURN_NOT_OK((launch_scatter<kAoS, 512, 16, 2, true, true>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow)));
      else
        B2_RETURN_NOT_OK((launch_scatter<kAoS, 512, 16, 2, true>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow)));
    } else {
      switch (ctx->tune[B2_TUNE_SCATTER_SHAPE]) {
        case 1: B2_RETURN_NOT_OK((launch_scatter<kAoS, 512, 8, 3>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow))); break;
        case 2: B2_RETURN_NOT_OK((launch_scatter<kAoS, 256, 16, 4>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow))); break;
        case 3: B2_RETURN_NOT_OK((launch_scatter<kAoS, 1024, 8, 2>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow))); break;
        case 8: B2_RETURN_NOT_OK((launch_scatter<kAoS, 1024, 16, 1>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow))); break;
        default:
          if (ctx->tune[B2_TUNE_SCATTER_PREFETCH])
            B2_RETURN_NOT_OK((launch_scatter<kAoS, 512, 16, 2, false, true>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow)));
          else
            B2_RETURN_NOT_OK((launch_scatter<kAoS, 512, 16, 2>(ctx, L.max_units, g.bits, s, in, d_seg_off, unit_first, nseg, L.unit_rows, g, scanned, d_out, out_cap, d_overflow)));
          break;
      }
    }
    B2_LAUNCH_CHECK(ctx, "part_scatter_kernel");
  }
  return B2_OK;
}

}  // namespace

size_t part_pass_ws_bytes(int64_t n, int64_t nseg, int bits) { return pass_layout(n, nseg, bits).total; }

int part_count(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
               const PartGeom& g, int64_t* d_part_off, void* d_ws, size_t ws_bytes, cudaStream_t s,
               int64_t clamp_rows) {
  B2_REQUIRE(ctx, g.bits >= 0 && g.bits <= kPartMaxBits, "fan-out per pass is limited to 2^10");
  B2_REQUIRE(ctx, nseg >= 1, "at least one segment");
  if (in.pairs)
    return part_count_impl<true>(ctx, in, n, d_seg_off, nseg, g, d_part_off, d_ws, ws_bytes, s, clamp_rows);
  return part_count_impl<false>(ctx, in, n, d_seg_off, nseg, g, d_part_off, d_ws, ws_bytes, s, clamp_rows);
}

int part_scatter(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
                 const PartGeom& g, uint2* d_out, int64_t out_cap, const uint64_t* d_bucket_addr,
                 unsigned int* d_overflow, void* d_ws, size_t ws_bytes, cudaStream_t s, const int64_t* d_abort) {
  if (in.pairs)
    return part_scatter_impl<true>(ctx, in, n, d_seg_off, nseg, g, d_out, out_cap, d_bucket_addr,
                                   d_overflow, d_ws, ws_bytes, s, d_abort);
  return part_scatter_impl<false>(ctx, in, n, d_seg_off, nseg, g, d_out, out_cap, d_bucket_addr,
                                  d_overflow, d_ws, ws_bytes, s, d_abort);
}

int part_pass(b2_ctx* ctx, const PartInput& in, int64_t n, const int64_t* d_seg_off, int64_t nseg,
              const PartGeom& g, uint2* d_out, int64_t out_cap, int64_t* d_part_off,
              unsigned int* d_overflow, void* d_ws, size_t ws_bytes, cudaStream_t s) {
  B2_RETURN_NOT_OK(part_count(ctx, in, n, d_seg_off, nseg, g, d_part_off, d_ws, ws_bytes, s, out_cap));
  return part_scatter(ctx, in, n, d_seg_off, nseg, g, d_out, out_cap, nullptr, d_overflow, d_ws,
                      ws_bytes, s);
}

// ---- one- or two-pass driver ------------------------------------------------------------------
namespace {

__global__ void set_single_segment_kernel(int64_t* seg_off, int64_t n) {
  seg_off[0] = 0;
  seg_off[1] = n;
}

struct FullLayout {
  int bits1, bits2;
  size_t off_seg, off_off1, off_pass, pass_bytes, total;
};

FullLayout full_layout(int64_t n, int bits) {
  FullLayout F;
  F.bits1 = bits <= kPartMaxBits ? bits : (bits + 1) / 2;
  F.bits2 = bits - F.bits1;
  size_t o = 0;
  F.off_seg = o;  o += 256;
  F.off_off1 = o; o += b2_align_up((((size_t)1 << F.bits1) + 1) * 8, 256);
  F.pass_bytes = part_pass_ws_bytes(n, 1, F.bits1);
  if (F.bits2 > 0) {
    const size_t p2 = part_pass_ws_bytes(n, (int64_t)1 << F.bits1, F.bits2);
    if (p2 > F.pass_bytes) F.pass_bytes = p2;
  }
  F.off_pass = o; o += b2_align_up(F.pass_bytes, 256);
  F.total = o;
  return F;
}

}  // namespace

size_t part_full_ws_bytes(int64_t n, int bits) { return full_layout(n, bits).total; }

int part_full(b2_ctx* ctx, const PartInput& in, int64_t n, int bits, int shl, int sel_shl,
              int sel_bits, uint32_t sel_val, uint2* d_out, uint2* d_tmp, int64_t cap,
              int64_t* d_off, unsigned int* d_overflow, void* d_ws, size_t ws_bytes,
              cudaStream_t s, bool val_pred, uint32_t val_thr) {
  B2_REQUIRE(ctx, !val_pred || in.pairs || in.vals, "a value predicate needs a value column");
  B2_REQUIRE(ctx, bits >= 0 && bits <= 2 * kPartMaxBits, "at most 2^20 partitions");
  const FullLayout F = full_layout(n, bits);
  if (ws_bytes < F.total) return b2_set_error(ctx, B2_ERR_WORKSPACE, "partition", "workspace");
  char* base = static_cast<char*>(d_ws);
  int64_t* seg = reinterpret_cast<int64_t*>(base + F.off_seg);
  int64_t* off1 = reinterpret_cast<int64_t*>(base + F.off_off1);
  set_single_segment_kernel<<<1, 1, 0, s>>>(seg, n);
  B2_LAUNCH_CHECK(ctx, "set_single_segment_kernel");
  PartGeom g1;
  g1.bits = F.bits1;
  g1.shl = shl;
  g1.sel_bits = sel_bits;
  g1.sel_shl = sel_shl;
  g1.sel_val = sel_val;
  g1.val_pred = val_pred;  // the first pass drops the rows; a second pass only sees survivors
  g1.val_thr = val_thr;
  if (F.bits2 == 0) {
    return part_pass(ctx, in, n, seg, 1, g1, d_out, cap, d_off, d_overflow, base + F.off_pass,
                     F.pass_bytes, s);
  }
  B2_REQUIRE(ctx, d_tmp != nullptr, "two-pass partitioning needs a temporary buffer");
  B2_RETURN_NOT_OK(part_pass(ctx, in, n, seg, 1, g1, d_tmp, cap, off1, d_overflow,
                             base + F.off_pass, F.pass_bytes, s));
  PartInput in2;
  in2.pairs = d_tmp;
  PartGeom g2;
  g2.bits = F.bits2;
  g2.shl = shl + F.bits1;
  // n is only used to size the work units (an upper bound on the rows that survived pass 1)
  return part_pass(ctx, in2, n, off1, (int64_t)1 << F.bits1, g2, d_out, cap, d_off,
                   d_overflow, base + F.off_pass, F.pass_bytes, s);
}

