// join.cu — inner equi-join L.fk = R.pk over uint32 keys with one uint32 payload per side.
//
// Replaces the reference's DPU join: JoinDpu::Run_internal (host/join/join_dpu.cc:168-400) moves
// every column four times across the DDR bus (partition, take, build/probe, take); on the DPU,
// kernel_hash_build (dpu/shared/kernels/hash_build.c:9-35) inserts key -> row index into an
// MRAM linear-probing table under 16 hardware mutexes (ht_put, dpu/shared/hashtable/
// hashtable.c:89-165, duplicate keys overwrite :133), kernel_hash_probe (hash_probe.c:9-46,
// ht_get hashtable.c:167-192) emits the matching row index and ASSERTS every probe hits
// (hash_probe.c:32-33), and a separate take pass gathers the payload (join_dpu.cc:303-368).
//
// Semantics here are Arrow's inner hash join (host/join/join_native.cc:31-36), which the
// reference's tests use as the oracle: unmatched probe rows are dropped, duplicate build keys
// produce one output row per duplicate. Output columns (fk, y, x); row order unspecified.
//
// B200 design:
//   1. both sides are radix-partitioned on the same hash bits (partition.cu) into partitions of
//      ~4096 build rows (up to 16384 on the perfect-hash path), carrying (key, payload) pairs — the
//      payload travels with the key, so there is no row-index vector and no take pass;
//   2. join_probe_kernel: one CTA per partition, table in SHARED memory, (fk, y, x) written with
//      coalesced stores, the output range of every 4608-row probe round reserved with one 64-bit
//      atomicAdd; the next partition's rows (and the next round's) are requested into L2 ahead of time.
//      Perfect-hash path (large joins: <= 14 hash bits left below the partition bits): the partition
//      hash is a bijection, so the remaining hash bits identify a key inside its partition — the
//      table is one value and one occupancy bit per entry, no keys, no collisions; an insert is a
//      store and an atom.or, a probe two loads and a bit test (see kDirectMaxBits below).
//      Bucketised path (small joins, and partitions with duplicate build keys): 2048 buckets x 4 slots,
//      72 KB with the values and one arrival counter per bucket (mean 2 rows per bucket). An insert is
//      one shared-memory atomicAdd on the counter — it returns the slot — plus two stores; a row goes
//      to the emptier of its TWO candidate buckets, so chains are rare; a probe reads both candidates
//      with 128-bit loads, straight-line. No empty-key marker: all 2^32 key values are legal, and only
//      the 8 KB of counters are cleared per partition. Build partitions larger than the table (skew,
//      heavy duplicates) are processed in chunks, each chunk probed by the whole probe partition.
//      kAgg variant: the fused [filter ->] join -> aggregate pipeline adds every output row's y and
//      x to per-thread sums instead of storing it (b2_join_aggr_u32_dev).
//   3. when the workspace is too small to hold both partitioned sides at once, the join runs in
//      hash-space slices (SF=2048 on one GPU fits in one: the output columns serve as the first
//      pass's temporary): slice s only partitions and joins the rows whose
//      top hash bits equal s.
#include <algorithm>

#include "partition.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBucketBits = 11;
constexpr int kBuckets = 1 << kBucketBits;     // 2048 buckets of 4 keys: one 128-bit read each
constexpr int kSlots = kBuckets * 4;           // 8192 slots = 32 KB keys + 32 KB values
constexpr int kTableBytes = (2 * kSlots + kBuckets) * 4;  // keys | values | per-bucket arrival counts
constexpr int kMaxBuild = (kSlots * 3) / 4;    // rows per build chunk (load factor <= 0.75)
constexpr int kTargetBuild = 4096;             // mean build rows per partition
constexpr int kItems = 9;                      // rows per thread per round (build and probe)
constexpr int kRound = kThreads * kItems;      // 4608 rows: mean partition + 8 sigma in one round
constexpr int kProbeCtasPerSm = 2;             // 3 CTAs/SM (40 registers, spills) measured 5 % slower

// Perfect-hash path of the probe kernel. wang_hash_u32 is a BIJECTION on 32-bit keys (odd
// multiplications and xor-shifts only), and every row of a partition has the same hash bits above the
// low `rest` ones (the GPU / slice / partition fields). So two DIFFERENT keys of one partition differ
// in the low `rest` bits of their hash: those bits index a table without collisions, whatever the key
// set is — and whoever sits in the entry a probe key maps to IS that key, so the table stores no keys
// at all: one 32-bit value per entry plus one occupancy bit. With rest <= 14 that is 64 KB + 2 KB, the
// shared memory the bucketised table uses, for partitions of up to 16384 distinct keys (four times the
// bucketised table's rows: both radix passes get away with a fan-out of 512 at SF=2048). An insert is
// one store and one shared-memory atomicOr on the occupancy word, a probe two 32-bit loads and a bit
// test: no candidate buckets, no arrival counters, no key compares, no chains (the bucketised probe
// loop issued ~98 lane-instructions per row and kept the shared-memory pipe 73 % busy,
// profiles/r2_join_probe.md). Only the 2 KB of occupancy bits are cleared per partition.
// Only EQUAL keys can meet in an entry: the atomicOr finds the bit set, the partition is then known to
// hold duplicate build keys and is redone by the bucketised path, which enumerates them.
// PRECONDITION: every row handed to the kernel shares the hash bits above `rest` with its partition.
// The partitioner guarantees it for the bits it consumed itself; bits skipped on the caller's word
// (hash_skip_bits) are only trusted where the library routed the rows (the segmented entry points
// behind the fused shuffle), see join_impl / join_seg_impl.
constexpr int kDirectMaxBits = 14;             // 16384 values x 4 B + 512 occupancy words = 66 KB <= kTableBytes
static_assert((4u << kDirectMaxBits) + (1u << kDirectMaxBits) / 8 <= (unsigned)kTableBytes,
              "the perfect-hash table shares the bucketised table's memory");

struct JoinState {  // lives in the workspace header
  unsigned long long out_rows;
  unsigned int overflow;
  unsigned int pad;
  unsigned long long sum_y, sum_x;  // fused join -> aggregate (b2_join_aggr_u32_dev)
};

// Fused pipeline [filter L.y < y_thr ->] join -> aggregate: nothing is materialised, the probe
// kernel adds every output row's y and x to per-thread sums instead of storing it.
struct JoinAggCfg {
  uint32_t y_thr = 0;
  bool filter_y = false;
  b2_join_aggr* d_out = nullptr;
};

// Bucket inside a partition's table. All keys of a partition share the TOP bits of wang_hash, so
// the table uses an independent multiplicative hash of the key itself (two instructions).
__device__ __forceinline__ uint32_t bucket_hash(uint32_t key) {
  return (key * 0x9E3779B1u) >> (32 - kBucketBits);
}
// Second candidate bucket (two-choice placement): another odd multiplier.
__device__ __forceinline__ uint32_t bucket_hash2(uint32_t key) {
  return (key * 0x85EBCA6Bu) >> (32 - kBucketBits);
}

// Rows [0, n) of src, item q of thread tid = row q*kThreads + tid; n < 2^31.
__device__ __forceinline__ void load_round(const uint2* __restrict__ src, uint32_t n, uint32_t tid,
                                           uint32_t (&k)[kItems], uint32_t (&v)[kItems]) {
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const uint32_t i = q * kThreads + tid;
    k[q] = 0;
    v[q] = 0;
    if (i < n) {
      const uint2 kv = ld_stream_v2(src + i);
      k[q] = kv.x;
      v[q] = kv.y;
    }
  }
}

// Slots of bucket `cur` that hold `key`, among its first min(n, 4) (the valid ones), as a 4-bit mask.
__device__ __forceinline__ uint32_t hit_mask(const uint4& cur, uint32_t key, uint32_t n) {
  const uint32_t hit = (cur.x == key ? 1u : 0u) | (cur.y == key ? 2u : 0u) | (cur.z == key ? 4u : 0u) |
                       (cur.w == key ? 8u : 0u);
  return hit & ((1u << min(n, 4u)) - 1u);
}

// Calls f(x) for the payload of every build row matching `key`, along the row's search path: first
// candidate bucket, second, then the chain behind the second while buckets overflowed.
template <typename F>
__device__ __forceinline__ void for_each_match(const uint32_t* cnt, const uint4* tk4, const uint32_t* tv,
                                               uint32_t key, F&& f) {
  const uint32_t h1 = bucket_hash(key), h2 = bucket_hash2(key);
  const bool chain = cnt[h1] > 4u && cnt[h2] > 4u;
  uint32_t b = h1;
  for (int step = 0;; ++step) {
    const uint32_t n = cnt[b];
    const uint4 cur = tk4[b];
    // from the second step on, the first candidate has been looked at already
    uint32_t hit = (step >= 1 && b == h1) ? 0u : hit_mask(cur, key, n);
    while (hit) {
      const int sidx = __ffs(hit) - 1;
      hit &= hit - 1;
      f(tv[b * 4 + sidx]);
    }
    if (step == 0) { b = h2; continue; }
    if (step == 1 ? !chain : n <= 4u) break;
    b = (b + 1) & (kBuckets - 1);
  }
}

// One CTA per partition; the table lives in shared memory.
//
// Table: 2048 buckets of 4 slots plus one ARRIVAL COUNTER per bucket. An insert is one shared-memory
// atomicAdd on the counter — the returned value is the slot, >= 4 sends the row on to the next
// bucket — and two stores: no compare-and-swap, no re-read of the bucket, no retry, no empty-key
// marker (so all 2^32 keys are legal and nothing but the 8 KB of counters is cleared per
// partition). A probe reads the counter and the bucket's four keys (one 128-bit load); only the
// first min(count, 4) slots are valid, and the chain continues in the next bucket only when
// count > 4, i.e. when a row really overflowed. (The first version found a free slot by reading
// the bucket and claiming it with atomicCAS, and probed until it met a bucket with a free slot:
// 86 + 131 lane-instructions per build + probe row, 40 % of them in per-item retry / chain loops
// that ran with 3-16 of 32 lanes active, profiles/r1_join.md.)
// TWO-CHOICE placement: a row has two candidate buckets (two multiplicative hashes) and goes to the
// emptier one, which keeps practically every bucket at <= 4 arrivals at a mean of 2 per bucket; a
// probe always looks at both candidates, straight-line, and only rows whose two candidates BOTH
// overflowed (count > 4) walk a chain (linear, from the second candidate). With one candidate
// bucket 5 % of the probe rows and 14 % of the build rows had to walk on, so with 288 rows per warp
// and round every per-item chain block ran ~2 times with ~3 of 32 lanes active: 75 + 40
// lane-instructions per row pair (profiles/r1_join.md).
// All global loads of a partition (its build rows and the first round of probe rows) are issued
// before anything else, so the clear and the inserts run under that latency.
// kAgg: the fused join -> aggregate pipeline (no output columns, no output-range reservation).
template <bool kAgg>
__global__ void __launch_bounds__(kThreads, kProbeCtasPerSm)
join_probe_kernel(const uint2* __restrict__ rpairs, const int64_t* __restrict__ roff,
                  const uint2* __restrict__ lpairs, const int64_t* __restrict__ loff,
                  int64_t nparts, uint32_t* __restrict__ out_fk, uint32_t* __restrict__ out_y,
                  uint32_t* __restrict__ out_x, int64_t out_cap, JoinState* __restrict__ st,
                  uint32_t y_thr, bool filter_y, int rest_bits) {
  extern __shared__ __align__(16) uint32_t tab[];  // keys [kSlots] | values [kSlots] | counts [kBuckets]
  uint32_t* tk = tab;
  uint32_t* tv = tab + kSlots;
  uint32_t* cnt = tab + 2 * kSlots;
  const uint4* tk4 = reinterpret_cast<const uint4*>(tab);
  __shared__ uint32_t warp_cnt[kWarps];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_dup;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = lanemask_lt();
  unsigned long long agg_rows = 0, agg_y = 0, agg_x = 0;  // kAgg: this thread's share of the aggregates

  // One round of probe rows whose match count is 0 or 1 (m[q]), in (warp, item, lane) order: the CTA
  // reserves its output range with one 64-bit atomic, ranks come from one ballot per item, positions
  // are 32-bit offsets from the warp's base, rows at or beyond out_cap are dropped.
  auto emit_unique = [&](const uint32_t (&lk)[kItems], const uint32_t (&ly)[kItems], const uint32_t (&x0)[kItems],
                         const uint32_t (&m)[kItems]) {
    uint32_t wtotal = 0;
#pragma unroll
    for (int q = 0; q < kItems; ++q) wtotal += __popc(__ballot_sync(0xffffffffu, m[q] == 1));
    if (lane == 0) warp_cnt[warp] = wtotal;
    __syncthreads();
    {
      const uint32_t w = lane < kWarps ? warp_cnt[lane] : 0;
      uint32_t incl = w;
#pragma unroll
      for (int o = 1; o < kWarps; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, kWarps - 1);
      wtotal = __shfl_sync(0xffffffffu, incl - w, warp);  // now: this warp's offset in the round
      if (tid == 0 && total > 0) s_base = atomicAdd(&st->out_rows, (unsigned long long)total);
    }
    __syncthreads();
    const unsigned long long pos = s_base + wtotal;
    const uint32_t room = (int64_t)pos < out_cap ? (uint32_t)min(out_cap - (int64_t)pos, (int64_t)kRound) : 0u;
    uint32_t* __restrict__ pf = out_fk + pos;
    uint32_t* __restrict__ py = out_y + pos;
    uint32_t* __restrict__ px = out_x + pos;
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < kItems; ++q) {
      const uint32_t one = __ballot_sync(0xffffffffu, m[q] == 1);
      const uint32_t off = run + __popc(one & lt);
      if (m[q] == 1 && off < room) {
        st_stream_u32(pf + off, lk[q]);
        st_stream_u32(py + off, ly[q]);
        st_stream_u32(px + off, x0[q]);
      }
      run += __popc(one);
    }
    // warp_cnt / s_base are rewritten only after the next round's first barrier
  };

  for (int64_t p = blockIdx.x; p < nparts; p += gridDim.x) {
    const int64_t r0 = roff[p], r1 = roff[p + 1];
    const int64_t l0 = loff[p], l1 = loff[p + 1];
    if (r1 == r0 || l1 == l0) continue;  // inner join: nothing to emit

    // The partition this CTA takes next is requested into L2 now (two copy-engine instructions by one
    // thread of the last warp): its loads, issued at the top of the next iteration, then hit L2 instead
    // of waiting for DRAM with only two CTAs per SM to hide it.
    if (tid == kThreads - 32 && p + gridDim.x < nparts) {
      const int64_t pn = p + gridDim.x;
      const int64_t nr0 = roff[pn], nr1 = roff[pn + 1], nl0 = loff[pn], nl1 = loff[pn + 1];
      if (nr1 > nr0 && nl1 > nl0) {
        l2_prefetch(rpairs + nr0, min(nr1 - nr0, (int64_t)kMaxBuild) * 8);
        l2_prefetch(lpairs + nl0, min(nl1 - nl0, (int64_t)kRound) * 8);
      }
    }
    if (rest_bits > 0) {
      // ---- perfect-hash path: entry = low rest_bits of wang_hash(key); values | occupancy bits ----
      const uint32_t mask = (1u << rest_bits) - 1u;
      uint32_t* occ = tab + (1u << rest_bits);
      const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab), occ_s = tab_s + (4u << rest_bits);
      uint32_t lk[kItems], ly[kItems];
      uint32_t dup = 0;
      // rounds after the first of either side are requested into L2 one round ahead (the partition-level
      // request above only covers the first round of each side)
      auto prefetch_round = [&](const uint2* base, int64_t from, int64_t end) {
        if (tid == kThreads - 64 && from < end) l2_prefetch(base + from, min(end - from, (int64_t)kRound) * 8);
      };
      {
        uint32_t rk[kItems], rv[kItems];
        const uint32_t nb0 = (uint32_t)min(r1 - r0, (int64_t)kRound);
        load_round(rpairs + r0, nb0, tid, rk, rv);
        load_round(lpairs + l0, (uint32_t)min(l1 - l0, (int64_t)kRound), tid, lk, ly);
        prefetch_round(rpairs, r0 + kRound, r1);
        __syncthreads();  // the previous partition is done with the table
        for (uint32_t i = tid; i < (1u << rest_bits) / 128; i += kThreads)
          reinterpret_cast<uint4*>(occ)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (rest_bits < 7 && tid < 4) occ[tid] = 0u;
        if (tid == 0) s_dup = 0;
        __syncthreads();
        for (int64_t base = r0; base < r1; base += kRound) {
          const uint32_t nb = (uint32_t)min(r1 - base, (int64_t)kRound);
          if (base > r0) load_round(rpairs + base, nb, tid, rk, rv);
          prefetch_round(rpairs, base + 2 * kRound, r1);
          if (base + kRound >= r1) prefetch_round(lpairs, l0 + kRound, l1);
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            if (nb == (uint32_t)kRound || q * kThreads + tid < nb) {
              const uint32_t idx = wang_hash_u32(rk[q]) & mask;
              const uint32_t bit = 1u << (idx & 31u);
              uint32_t old;
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_s + idx * 4u), "r"(rv[q]) : "memory");
              asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(occ_s + (idx >> 5) * 4u), "r"(bit) : "memory");
              dup |= old & bit;  // only an equal key can have been there
            }
          }
        }
        if (dup) s_dup = 1;
      }
      __syncthreads();
      if (s_dup == 0) {
        for (int64_t t0 = l0; t0 < l1; t0 += kRound) {
          const uint32_t nprobe = (uint32_t)min(l1 - t0, (int64_t)kRound);
          if (t0 > l0) load_round(lpairs + t0, nprobe, tid, lk, ly);
          if (t0 > l0) prefetch_round(lpairs, t0 + kRound, l1);
          uint32_t x0[kItems], m[kItems];
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            const uint32_t idx = wang_hash_u32(lk[q]) & mask;
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(occ_s + (idx >> 5) * 4u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x0[q]) : "r"(tab_s + idx * 4u));
            const bool active = (nprobe == (uint32_t)kRound || q * kThreads + tid < nprobe) &&
                                (!kAgg || !filter_y || ly[q] < y_thr);
            m[q] = active ? (w >> (idx & 31u)) & 1u : 0u;
          }
          if (kAgg) {
#pragma unroll
            for (int q = 0; q < kItems; ++q) {
              if (m[q]) {
                agg_x += x0[q];
                agg_y += ly[q];
                agg_rows += 1;
              }
            }
          } else {
            emit_unique(lk, ly, x0, m);
          }
        }
        continue;  // the next partition's first barrier separates its table writes from these reads
      }
      // duplicate build keys: the bucketised path below redoes this partition (nothing was emitted)
    }
    for (int64_t c0 = r0; c0 < r1; c0 += kMaxBuild) {
      const uint32_t nbuild = (uint32_t)min((int64_t)kMaxBuild, r1 - c0);
      uint32_t lk[kItems], ly[kItems];
      {
        uint32_t rk[kItems], rv[kItems];
        load_round(rpairs + c0, nbuild, tid, rk, rv);
        load_round(lpairs + l0, (uint32_t)min(l1 - l0, (int64_t)kRound), tid, lk, ly);

        __syncthreads();  // the previous probe phase is done with the table
        if (tid < kBuckets / 4) reinterpret_cast<uint4*>(cnt)[tid] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (uint32_t base = 0; base < nbuild; base += kRound) {
          if (base > 0) load_round(rpairs + c0 + base, nbuild - base, tid, rk, rv);
          // e = stage << 16 | bucket * 8 + slot; slot 4 = "bucket was full, still looking";
          // stage 0 = first candidate tried, 1 = second, 2 = walking the chain behind the second
          uint32_t e[kItems];
          uint32_t pend = 0;
#pragma unroll
          for (int q = 0; q < kItems; ++q) {  // all first-choice counters are bumped back to back
            const uint32_t h1 = bucket_hash(rk[q]), h2 = bucket_hash2(rk[q]);
            const uint32_t b = cnt[h2] < cnt[h1] ? h2 : h1;  // racy read: only a placement heuristic
            e[q] = b * 8;
            if (base + q * kThreads + tid < nbuild) {
              e[q] += min(atomicAdd(&cnt[b], 1u), 4u);
              pend |= ((e[q] & 7u) == 4u ? 1u : 0u) << q;
            }
          }
          while (pend) {  // rare: both candidates full
#pragma unroll
            for (int q = 0; q < kItems; ++q) {
              if (pend >> q & 1) {
                const uint32_t stage = e[q] >> 16, was = (e[q] & 0xffffu) >> 3;
                const uint32_t h1 = bucket_hash(rk[q]), h2 = bucket_hash2(rk[q]);
                const uint32_t b = stage == 0 ? (was ^ h1 ^ h2) : (stage == 1 ? h2 + 1 : was + 1) & (kBuckets - 1);
                e[q] = (min(stage + 1, 2u) << 16) | (b * 8 + min(atomicAdd(&cnt[b], 1u), 4u));
                if ((e[q] & 7u) != 4u) pend &= ~(1u << q);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            if (base + q * kThreads + tid < nbuild) {
              const uint32_t slot = ((e[q] & 0xffffu) >> 3) * 4 + (e[q] & 7u);
              tk[slot] = rk[q];
              tv[slot] = rv[q];  // duplicates of a key take separate slots
            }
          }
        }
      }
      __syncthreads();

      for (int64_t t0 = l0; t0 < l1; t0 += kRound) {
        const uint32_t nprobe = (uint32_t)min(l1 - t0, (int64_t)kRound);
        if (t0 > l0) load_round(lpairs + t0, nprobe, tid, lk, ly);
        uint32_t x0[kItems], m[kItems];
        uint32_t pend = 0;
#pragma unroll
        for (int q = 0; q < kItems; ++q) {  // both candidate buckets of every item: straight-line
          const uint32_t h1 = bucket_hash(lk[q]), h2 = bucket_hash2(lk[q]);
          const uint32_t n1 = cnt[h1], n2 = cnt[h2];
          const uint4 c1 = tk4[h1], c2 = tk4[h2];
          const bool active = q * kThreads + tid < nprobe && (!kAgg || !filter_y || ly[q] < y_thr);
          const uint32_t hit1 = active ? hit_mask(c1, lk[q], n1) : 0u;
          const uint32_t hit2 = (active && h2 != h1) ? hit_mask(c2, lk[q], n2) : 0u;
          m[q] = __popc(hit1) + __popc(hit2);
          x0[q] = 0;
          if (hit1) x0[q] = tv[h1 * 4 + (__ffs(hit1) - 1)];
          else if (hit2) x0[q] = tv[h2 * 4 + (__ffs(hit2) - 1)];
          pend |= (active && n1 > 4u && n2 > 4u ? 1u : 0u) << q;  // rows overflowed out of both
        }
        if (pend) {  // rare: the chain behind the second candidate
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            if (pend >> q & 1) {
              const uint32_t h1 = bucket_hash(lk[q]);
              uint32_t b = bucket_hash2(lk[q]);
              uint32_t n;
              do {
                b = (b + 1) & (kBuckets - 1);
                n = cnt[b];
                const uint4 cur = tk4[b];
                // the chain may run through the first candidate, which has been looked at already
                const uint32_t hit = b == h1 ? 0u : hit_mask(cur, lk[q], n);
                if (hit) {
                  if (m[q] == 0) x0[q] = tv[b * 4 + (__ffs(hit) - 1)];
                  m[q] += __popc(hit);
                }
              } while (n > 4u);
            }
          }
        }
        if (kAgg) {  // every output row (fk, y, x) adds y and x to the sums; nothing is stored
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            if (m[q] == 1) {
              agg_x += x0[q];
            } else if (m[q] > 1) {
              for_each_match(cnt, tk4, tv, lk[q], [&](uint32_t x) { agg_x += x; });
            }
            agg_rows += m[q];
            agg_y += (unsigned long long)ly[q] * m[q];
          }
          continue;
        }
        // ---- output positions in (warp, item, lane) order: one 64-bit atomic per CTA and round ----
        uint32_t wtotal = 0;
        uint32_t many = 0;
#pragma unroll
        for (int q = 0; q < kItems; ++q) many |= m[q];
        // duplicate build keys anywhere in this warp's round? (one vote instead of one per item)
        const bool warp_multi = __any_sync(0xffffffffu, many > 1u);
        if (!warp_multi) {
#pragma unroll
          for (int q = 0; q < kItems; ++q) wtotal += __popc(__ballot_sync(0xffffffffu, m[q] == 1));
        } else {
#pragma unroll
          for (int q = 0; q < kItems; ++q) wtotal += __reduce_add_sync(0xffffffffu, m[q]);
        }
        if (lane == 0) warp_cnt[warp] = wtotal;
        __syncthreads();
        {
          // every warp scans the 16 warp totals itself; warp 0 reserves the CTA's output range
          const uint32_t w = lane < kWarps ? warp_cnt[lane] : 0;
          uint32_t incl = w;
#pragma unroll
          for (int o = 1; o < kWarps; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const uint32_t total = __shfl_sync(0xffffffffu, incl, kWarps - 1);
          wtotal = __shfl_sync(0xffffffffu, incl - w, warp);  // now: this warp's offset in the round
          if (tid == 0 && total > 0) s_base = atomicAdd(&st->out_rows, (unsigned long long)total);
        }
        __syncthreads();
        unsigned long long pos = s_base + wtotal;
        if (!warp_multi) {
          // unique build keys (the common case): ranks come from one ballot per item, positions are
          // 32-bit offsets from the warp's base, rows at or beyond out_cap are dropped
          const uint32_t room = (int64_t)pos < out_cap ? (uint32_t)min(out_cap - (int64_t)pos, (int64_t)kRound) : 0u;
          uint32_t* __restrict__ pf = out_fk + pos;
          uint32_t* __restrict__ py = out_y + pos;
          uint32_t* __restrict__ px = out_x + pos;
          uint32_t run = 0;
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            const uint32_t one = __ballot_sync(0xffffffffu, m[q] == 1);
            const uint32_t off = run + __popc(one & lt);
            if (m[q] == 1 && off < room) {
              st_stream_u32(pf + off, lk[q]);
              st_stream_u32(py + off, ly[q]);
              st_stream_u32(px + off, x0[q]);
            }
            run += __popc(one);
          }
        } else {
          // duplicate build keys somewhere in this warp's round: enumerate every match
#pragma unroll
          for (int q = 0; q < kItems; ++q) {
            uint32_t incl = m[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            unsigned long long pp = pos + incl - m[q];
            pos += __shfl_sync(0xffffffffu, incl, 31);
            if (m[q] == 1) {
              if ((int64_t)pp < out_cap) {
                out_fk[pp] = lk[q];
                out_y[pp] = ly[q];
                out_x[pp] = x0[q];
              }
            } else if (m[q] > 1) {
              for_each_match(cnt, tk4, tv, lk[q], [&](uint32_t x) {
                if ((int64_t)pp < out_cap) {
                  out_fk[pp] = lk[q];
                  out_y[pp] = ly[q];
                  out_x[pp] = x;
                }
                ++pp;
              });
            }
          }
        }
        // warp_cnt / s_base are rewritten only after the next round's first barrier
      }
    }
    __syncthreads();
  }
  if (kAgg) {
    agg_rows = warp_reduce_sum_u64(agg_rows);
    agg_y = warp_reduce_sum_u64(agg_y);
    agg_x = warp_reduce_sum_u64(agg_x);
    if (lane == 0 && agg_rows) {
      atomicAdd(&st->out_rows, agg_rows);
      atomicAdd(&st->sum_y, agg_y);
      atomicAdd(&st->sum_x, agg_x);
    }
  }
}

__global__ void join_init_kernel(JoinState* st) {
  st->out_rows = 0;
  st->overflow = 0;
  st->sum_y = 0;
  st->sum_x = 0;
}
__global__ void join_aggr_finish_kernel(const JoinState* __restrict__ st, b2_join_aggr* __restrict__ out,
                                        const int64_t* __restrict__ abort_flag = nullptr) {
  // a partition buffer that overflowed (skewed slice) or an exchange that was called off on the device
  // makes the result invalid: report ~0 rows
  out->rows = (st->overflow || (abort_flag && *abort_flag)) ? ~0ull : st->out_rows;
  out->sum_y = st->sum_y;
  out->sum_x = st->sum_x;
}
__global__ void join_finish_kernel(const JoinState* __restrict__ st, uint64_t* __restrict__ out_rows,
                                   const int64_t* __restrict__ abort_flag = nullptr) {
  // a partition buffer that overflowed (skewed slice) makes the result invalid: report ~0; so does
  // an exchange that was called off on the device (receive capacity exceeded, b2_shuffle_p2p_plan_dev)
  *out_rows = (st->overflow || (abort_flag && *abort_flag)) ? ~0ull : st->out_rows;
}

int ceil_log2_i64(int64_t v) {
  int b = 0;
  while (((int64_t)1 << b) < v) ++b;
  return b;
}

// Entry bits of the perfect-hash path for a partitioning that consumed `used` hash bits (0 = the
// bucketised table: too many bits left for a shared-memory table).
int direct_rest_bits(int used) {
  const int rest = 32 - used;
  return rest >= 1 && rest <= kDirectMaxBits ? rest : 0;
}

struct JoinPlan {
  int bits;        // fine partition bits
  int rest_bits;   // perfect-hash entry bits (0: bucketised table)
  int slice_bits;  // log2 number of hash-space slices
  int64_t cap_r, cap_l, cap_tmp;
  bool two_pass;
  size_t off_state, off_roff, off_loff, off_rout, off_lout, off_tmp, off_part, part_bytes, total;
};

// ext_tmp: the temporary of the first radix pass lives outside the workspace (in the caller's
// output columns, see join_impl).
JoinPlan make_plan(int64_t nl, int64_t nr, int skip_bits, int slice_bits, bool ext_tmp = false,
                   int direct_min_rows = 2048) {
  JoinPlan P;
  P.slice_bits = slice_bits;
  const int64_t nslices = (int64_t)1 << slice_bits;
  const int64_t nr_slice = (nr + nslices - 1) / nslices;
  int bits = ceil_log2_i64((nr_slice + kTargetBuild - 1) / kTargetBuild);
  bits = std::max(bits, 1);
  bits = std::min(bits, 2 * kPartMaxBits);
  bits = std::min(bits, 32 - skip_bits - slice_bits);
  // a few more partition bits than the table size asks for make the perfect-hash path possible
  // (<= 13 hash bits left), as long as partitions keep >= 1024 build rows
  const int direct_bits = 32 - skip_bits - slice_bits - kDirectMaxBits;
  if (direct_min_rows > 0 && skip_bits == 0 && direct_bits > bits && direct_bits <= 2 * kPartMaxBits &&
      (nr_slice >> direct_bits) >= direct_min_rows)
    bits = direct_bits;
  // ... and fewer bits than the bucketised table would want are enough when the perfect-hash table
  // (2^14 entries) still holds a partition's expected rows: both radix passes get cheaper
  if (direct_min_rows > 0 && skip_bits == 0 && direct_bits >= 1 && direct_bits < bits &&
      ((nr_slice + ((int64_t)1 << direct_bits) - 1) >> direct_bits) <= ((int64_t)1 << kDirectMaxBits))
    bits = direct_bits;
  P.bits = std::max(bits, 1);
  // bits skipped on the caller's word are not trusted (see the precondition of the perfect-hash path)
  P.rest_bits = direct_min_rows > 0 && skip_bits == 0 ? direct_rest_bits(slice_bits + P.bits) : 0;
  P.two_pass = P.bits > kPartMaxBits;
  // slices are hash-uniform in expectation; leave 12.5 % + 64 Ki rows of slack for skew
  auto cap = [&](int64_t n) {
    if (slice_bits == 0) return n;
    const int64_t per = (n + nslices - 1) / nslices;
    return std::min(n, per + per / 8 + 65536);
  };
  P.cap_r = cap(nr);
  P.cap_l = cap(nl);
  P.cap_tmp = P.two_pass ? std::max(P.cap_r, P.cap_l) : 0;
  const size_t tmp_bytes = ext_tmp ? 0 : (size_t)P.cap_tmp * 8;
  const size_t noff = (((size_t)1 << P.bits) + 1) * 8;
  size_t o = 0;
  P.off_state = o; o += 256;
  P.off_roff = o;  o += b2_align_up(noff, 256);
  P.off_loff = o;  o += b2_align_up(noff, 256);
  P.off_rout = o;  o += b2_align_up((size_t)P.cap_r * 8, 256);
  P.off_lout = o;  o += b2_align_up((size_t)P.cap_l * 8, 256);
  P.off_tmp = o;   o += b2_align_up(tmp_bytes, 256);
  P.part_bytes = std::max(part_full_ws_bytes(nr, P.bits), part_full_ws_bytes(nl, P.bits));
  P.off_part = o;  o += b2_align_up(P.part_bytes, 256);
  P.total = o;
  return P;
}

constexpr int kMaxSliceBits = 6;

int join_impl(b2_ctx* ctx, const PartInput& lin, int64_t nl, const PartInput& rin, int64_t nr,
              uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
              uint64_t* d_out_rows, int skip_bits, void* d_ws, size_t ws_bytes, cudaStream_t s,
              const JoinAggCfg* agg = nullptr, int phases = 3) {
  // phases: 1 = reset + the build side's radix passes (only rin has to be there), 2 = the probe side's
  // passes, the probe and the result; 3 = both. A sliced join (workspace-bound) does all its work in
  // phase 2, slice by slice.
  B2_REQUIRE(ctx, phases >= 1 && phases <= 3, "phases: 1 build | 2 probe");
  B2_REQUIRE(ctx, nl >= 0 && nr >= 0 && out_capacity >= 0, "negative size");
  B2_REQUIRE(ctx, skip_bits >= 0 && skip_bits <= 8, "hash_skip_bits must be in 0..8");
  B2_REQUIRE(ctx, agg ? agg->d_out != nullptr : d_out_rows != nullptr, "result pointer is null");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
             "workspace must be 256 B aligned");
  B2_REQUIRE(ctx, out_capacity == 0 || (d_out_fk && d_out_y && d_out_x), "null output column");
  // Output columns that are ONE allocation (fk | y | x back to back, 32 B aligned) are dead until the
  // probe kernel writes them, so they serve as the temporary of the first radix pass — the
  // reference aliases its outputs onto the partitioned left side in the same spirit
  // (join_dpu.cc:315-322). At SF=2048 on one GPU this is what lets the join run in ONE hash-space
  // slice: inputs (64 GiB) + outputs (48 GiB) + both partitioned sides (64 GiB) fit, a 32 GiB
  // temporary on top would not.
  const bool out_adjacent = !agg && out_capacity > 0 && d_out_y == d_out_fk + out_capacity &&
                            d_out_x == d_out_y + out_capacity &&
                            (reinterpret_cast<uintptr_t>(d_out_fk) & 31) == 0 &&
                            12 * (uint64_t)out_capacity >= 8 * (uint64_t)std::max(nl, nr);
  // pick the smallest number of slices whose plan fits the workspace
  const int dmin = ctx->tune[B2_TUNE_JOIN_DIRECT_MIN_ROWS];
  JoinPlan P = make_plan(nl, nr, skip_bits, 0, false, dmin);
  int sb = 0;
  bool ext_tmp = false;
  if (P.total > ws_bytes && out_adjacent && P.two_pass) {
    const JoinPlan Q = make_plan(nl, nr, skip_bits, 0, true, dmin);
    if (Q.total <= ws_bytes) {
      P = Q;
      ext_tmp = true;
    }
  }
  while (P.total > ws_bytes && sb < kMaxSliceBits) P = make_plan(nl, nr, skip_bits, ++sb, false, dmin);
  if (P.total > ws_bytes)
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "join workspace", "see b2_join_min_ws_bytes()");
  char* base = static_cast<char*>(d_ws);
  JoinState* st = reinterpret_cast<JoinState*>(base + P.off_state);
  int64_t* roff = reinterpret_cast<int64_t*>(base + P.off_roff);
  int64_t* loff = reinterpret_cast<int64_t*>(base + P.off_loff);
  uint2* rout = reinterpret_cast<uint2*>(base + P.off_rout);
  uint2* lout = reinterpret_cast<uint2*>(base + P.off_lout);
  uint2* tmp = P.two_pass ? reinterpret_cast<uint2*>(ext_tmp ? (char*)d_out_fk : base + P.off_tmp) : nullptr;
  void* pws = base + P.off_part;

  if (phases & 1) {
    if (ctx->trace_join) b2_trace_reset(ctx);
    join_init_kernel<<<1, 1, 0, s>>>(st);
    B2_LAUNCH_CHECK(ctx, "join_init_kernel");
  }
  if (nl > 0 && nr > 0) {
    static const int seen = b2_new_site();
    if (b2_first_use_on_device(ctx, seen)) {
      B2_CUDA_OK(ctx, cudaFuncSetAttribute(join_probe_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes));
      B2_CUDA_OK(ctx, cudaFuncSetAttribute(join_probe_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kTableBytes));
    }
    const int64_t nparts = (int64_t)1 << P.bits;
    const int part_shl = skip_bits + P.slice_bits;
    for (uint32_t slice = 0; slice < (1u << P.slice_bits); ++slice) {
      if (P.slice_bits == 0 ? (phases & 1) : (phases & 2)) {
        b2_trace_scope tr(ctx, B2_PHASE_PART_BUILD, s);
        B2_RETURN_NOT_OK(part_full(ctx, rin, nr, P.bits, part_shl, skip_bits, P.slice_bits, slice,
                                   rout, tmp, P.cap_r, roff, &st->overflow, pws, P.part_bytes, s));
      }
      if (!(phases & 2)) break;
      {  // filter pushdown: the probe side's first radix pass drops the rows that fail L.y < y_thr
        b2_trace_scope tr(ctx, B2_PHASE_PART_PROBE, s);
        B2_RETURN_NOT_OK(part_full(ctx, lin, nl, P.bits, part_shl, skip_bits, P.slice_bits, slice,
                                   lout, tmp, P.cap_l, loff, &st->overflow, pws, P.part_bytes, s,
                                   agg && agg->filter_y, agg ? agg->y_thr : 0u));
      }
      b2_trace_scope tr(ctx, B2_PHASE_PROBE, s);
      int64_t grid = std::min<int64_t>(nparts, (int64_t)ctx->sm_count * kProbeCtasPerSm);
      if (agg)
        join_probe_kernel<true><<<(unsigned)grid, kThreads, kTableBytes, s>>>(
            rout, roff, lout, loff, nparts, nullptr, nullptr, nullptr, 0, st, agg->y_thr, false, P.rest_bits);
      else
        join_probe_kernel<false><<<(unsigned)grid, kThreads, kTableBytes, s>>>(
            rout, roff, lout, loff, nparts, d_out_fk, d_out_y, d_out_x, out_capacity, st, 0u, false, P.rest_bits);
      B2_LAUNCH_CHECK(ctx, "join_probe_kernel");
    }
  }
  if (!(phases & 2)) return B2_OK;
  if (agg) {
    join_aggr_finish_kernel<<<1, 1, 0, s>>>(st, agg->d_out);
  } else {
    join_finish_kernel<<<1, 1, 0, s>>>(st, d_out_rows);
  }
  B2_LAUNCH_CHECK(ctx, "join_finish_kernel");
  return B2_OK;
}

// ---- standalone partition: permutation + gather (PartitionDpu) ------------------------------
__global__ void __launch_bounds__(256)
part_gather_kernel(const uint2* __restrict__ pairs, int64_t n, const uint32_t* __restrict__ col_in,
                   uint32_t* __restrict__ col_out, int take_key) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const uint2 kv = pairs[i];
    col_out[i] = take_key ? kv.x : __ldg(col_in + kv.y);
  }
}

// ---- join of pre-partitioned sides (the fused multi-GPU shuffle delivers them) -------------------
// Both sides arrive grouped into 2^seg_bits coarse buckets on hash bits [skip, skip + seg_bits),
// bucket boundaries in d_*_seg_off. Only the fine pass is left (or nothing at all when the coarse
// buckets are already table-sized).
struct SegPlan {
  int total_bits, fine_bits, rest_bits;
  size_t off_state, off_roff, off_loff, off_rout, off_lout, off_part, part_bytes, total;
};

// nl / nr size the buffers and work units (upper bounds are fine: the real row counts are the last
// entries of the segment tables on the device); nr_expected (<= 0: nr) picks the number of fine
// partitions, so a caller that only knows a generous receive CAPACITY still gets ~4096-row tables.
bool make_seg_plan(int64_t nl, int64_t nr, int skip_bits, int seg_bits, SegPlan* P, int64_t nr_expected = 0,
                   bool direct = true) {
  const int64_t nr_plan = nr_expected > 0 ? std::min(nr_expected, nr) : nr;
  int bits = ceil_log2_i64((nr_plan + kTargetBuild - 1) / kTargetBuild);
  bits = std::max(bits, seg_bits);
  bits = std::min(bits, 32 - skip_bits);
  // one fine pass refines a coarse bucket at most 2^10-fold; beyond that partitions simply get
  // larger than the table and the probe kernel builds them in chunks
  bits = std::min(bits, seg_bits + kPartMaxBits);
  // the perfect-hash table holds 2^14 distinct keys: fewer, larger partitions when it applies
  const int direct_bits = 32 - skip_bits - kDirectMaxBits;
  if (direct && direct_bits >= seg_bits && direct_bits < bits &&
      ((nr_plan + ((int64_t)1 << direct_bits) - 1) >> direct_bits) <= ((int64_t)1 << kDirectMaxBits))
    bits = direct_bits;
  P->total_bits = bits;
  P->rest_bits = direct ? direct_rest_bits(skip_bits + bits) : 0;
  P->fine_bits = bits - seg_bits;
  const size_t noff = (((size_t)1 << bits) + 1) * 8;
  size_t o = 0;
  P->off_state = o; o += 256;
  P->off_roff = o;  o += b2_align_up(noff, 256);
  P->off_loff = o;  o += b2_align_up(noff, 256);
  const bool fine = P->fine_bits > 0;
  P->off_rout = o;  o += fine ? b2_align_up((size_t)nr * 8, 256) : 0;
  P->off_lout = o;  o += fine ? b2_align_up((size_t)nl * 8, 256) : 0;
  const int64_t nseg = (int64_t)1 << seg_bits;
  P->part_bytes = fine ? std::max(part_pass_ws_bytes(nr, nseg, P->fine_bits),
                                  part_pass_ws_bytes(nl, nseg, P->fine_bits)) : 0;
  P->off_part = o;  o += b2_align_up(P->part_bytes, 256);
  P->total = o;
  return true;
}

int join_seg_impl(b2_ctx* ctx, const uint2* lpairs, const int64_t* l_seg_off, int64_t nl,
                  const uint2* rpairs, const int64_t* r_seg_off, int64_t nr, int seg_bits,
                  uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                  uint64_t* d_out_rows, int skip_bits, void* d_ws, size_t ws_bytes, cudaStream_t s,
                  int64_t nr_expected = 0, const int64_t* d_abort = nullptr, int phases = 7,
                  const JoinAggCfg* agg = nullptr) {
  B2_REQUIRE(ctx, nl >= 0 && nr >= 0 && out_capacity >= 0, "negative size");
  B2_REQUIRE(ctx, seg_bits >= 0 && seg_bits <= kPartMaxBits && skip_bits >= 0 && skip_bits + seg_bits <= 20,
             "bad skip/segment bits");
  B2_REQUIRE(ctx, r_seg_off && (l_seg_off || !(phases & 2)) && (d_out_rows || agg || !(phases & 4)), "null pointer");
  B2_REQUIRE(ctx, !agg || agg->d_out, "null aggregate result");
  B2_REQUIRE(ctx, phases >= 1 && phases <= 7, "phases: 1 build | 2 probe | 4 finish");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
             "workspace must be 256 B aligned");
  SegPlan P;
  // the rows were routed here by their top skip_bits hash bits and grouped on the next seg_bits (the
  // entry point's contract), so the perfect-hash path may rely on them
  if (!make_seg_plan(nl, nr, skip_bits, seg_bits, &P, nr_expected, ctx->tune[B2_TUNE_JOIN_DIRECT_MIN_ROWS] > 0))
    return b2_set_error(ctx, B2_ERR_UNSUPPORTED, "segmented join",
                        "build side too large for one fine pass; use b2_join_pairs_dev");
  if (P.total > ws_bytes)
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "segmented join workspace", "see b2_join_seg_ws_bytes()");
  char* base = static_cast<char*>(d_ws);
  JoinState* st = reinterpret_cast<JoinState*>(base + P.off_state);
  if (phases & 1) {
    if (ctx->trace_join) b2_trace_reset(ctx);
    join_init_kernel<<<1, 1, 0, s>>>(st);
    B2_LAUNCH_CHECK(ctx, "join_init_kernel");
  }
  if ((nl > 0 || !(phases & 2)) && nr > 0) {
    static const int seen = b2_new_site();
    if (b2_first_use_on_device(ctx, seen)) {
      B2_CUDA_OK(ctx, cudaFuncSetAttribute(join_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kTableBytes));
      B2_CUDA_OK(ctx, cudaFuncSetAttribute(join_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kTableBytes));
    }
    const int64_t nseg = (int64_t)1 << seg_bits;
    const int64_t nparts = (int64_t)1 << P.total_bits;
    const uint2 *rp = rpairs, *lp = lpairs;
    const int64_t *roff = r_seg_off, *loff = l_seg_off;
    if (P.fine_bits > 0) {
      PartGeom g;
      g.bits = P.fine_bits;
      g.shl = skip_bits + seg_bits;
      PartInput rin, lin;
      rin.pairs = rpairs;
      lin.pairs = lpairs;
      uint2* rout = reinterpret_cast<uint2*>(base + P.off_rout);
      uint2* lout = reinterpret_cast<uint2*>(base + P.off_lout);
      int64_t* roff_w = reinterpret_cast<int64_t*>(base + P.off_roff);
      int64_t* loff_w = reinterpret_cast<int64_t*>(base + P.off_loff);
      void* pws = base + P.off_part;
      // build phase: the build side's fine pass; its result stays in the workspace for every probe phase
      if (phases & 1) {
        b2_trace_scope tr(ctx, B2_PHASE_PART_BUILD, s);
        B2_RETURN_NOT_OK(part_pass(ctx, rin, nr, r_seg_off, nseg, g, rout, nr, roff_w, &st->overflow, pws,
                                   P.part_bytes, s));
      }
      if (phases & 2) {
        b2_trace_scope tr(ctx, B2_PHASE_PART_PROBE, s);
        B2_RETURN_NOT_OK(part_pass(ctx, lin, nl, l_seg_off, nseg, g, lout, nl, loff_w, &st->overflow, pws,
                                   P.part_bytes, s));
      }
      rp = rout; lp = lout; roff = roff_w; loff = loff_w;
    }
    if (phases & 2) {  // probe phase: may run several times, each over another share of the probe side
      b2_trace_scope tr(ctx, B2_PHASE_PROBE, s);
      const int64_t grid = std::min<int64_t>(nparts, (int64_t)ctx->sm_count * kProbeCtasPerSm);
      if (agg)  // the rows crossed the link unfiltered: the probe kernel evaluates L.y < y_thr itself
        join_probe_kernel<true><<<(unsigned)grid, kThreads, kTableBytes, s>>>(
            rp, roff, lp, loff, nparts, nullptr, nullptr, nullptr, 0, st, agg->y_thr, agg->filter_y, P.rest_bits);
      else
        join_probe_kernel<false><<<(unsigned)grid, kThreads, kTableBytes, s>>>(
            rp, roff, lp, loff, nparts, d_out_fk, d_out_y, d_out_x, out_capacity, st, 0u, false,
            P.rest_bits);
      B2_LAUNCH_CHECK(ctx, "join_probe_kernel");
    }
  }
  if (phases & 4) {
    if (agg) join_aggr_finish_kernel<<<1, 1, 0, s>>>(st, agg->d_out, d_abort);
    else join_finish_kernel<<<1, 1, 0, s>>>(st, d_out_rows, d_abort);
    B2_LAUNCH_CHECK(ctx, "join_finish_kernel");
  }
  return B2_OK;
}


// ---- fused shuffle: every rank's bucket boundaries -> this rank's destination addresses --------------
// One CTA, one thread per bucket (2^bits <= 1024). What dpu_olap_b200/sharded.py::p2p_plan did with a
// dozen eager torch kernels and a host read-back is one launch here, and the capacity check stays
// on the device.
__global__ void __launch_bounds__(1024)
shuffle_plan_kernel(const int64_t* const* __restrict__ off_ptrs, const uint64_t* __restrict__ recv_base, int rank,
                    int nranks, int bits, int64_t capacity_rows, uint64_t* __restrict__ bucket_addr,
                    int64_t* __restrict__ seg_off, int64_t* __restrict__ info,
                    const int64_t* __restrict__ prev_abort) {
  __shared__ int64_t excl[1025];
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t s_max;
  const int B = 1 << bits;
  const int C = B / nranks;  // coarse buckets per destination rank
  const int b = threadIdx.x;
  const uint32_t lane = b & 31, warp = b >> 5;
  int64_t tot = 0, before = 0;  // rows of all ranks / of the lower ranks in bucket b
  if (b < B) {
    for (int s = 0; s < nranks; ++s) {
      const int64_t* o = off_ptrs[s];
      const int64_t c = o[b + 1] - o[b];
      if (s < rank) before += c;
      tot += c;
    }
  }
  int64_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  if (b == 0) s_max = 0;
  __syncthreads();
  if (warp == 0) {
    const int64_t w = warp_tot[lane];
    int64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;
  }
  __syncthreads();
  const int64_t ex = warp_tot[warp] + incl - tot;
  if (b < B) excl[b] = ex;
  if (b == B - 1) excl[B] = ex + tot;
  __syncthreads();
  // rows every destination receives; the largest decides whether the exchange may run at all
  if (b < nranks) atomicMax(reinterpret_cast<unsigned long long*>(&s_max),
                            (unsigned long long)(excl[(b + 1) * C] - excl[b * C]));
  __syncthreads();
  // prev_abort chains the two sides of a join: the second plan's flag covers both
  const bool overflow = s_max > capacity_rows || (prev_abort && *prev_abort);
  if (b < B) {
    const int dest = b / C;
    // receive buffers are laid out bucket-major, source-minor: every coarse bucket is contiguous
    bucket_addr[b] = recv_base[dest] + 8ull * (uint64_t)(excl[b] - excl[dest * C] + before);
  }
  if (b < C) seg_off[b] = overflow ? 0 : excl[rank * C + b] - excl[rank * C];
  if (b == 0) {  // C can be all 1024 buckets (one rank): the closing boundary has no thread of its own
    seg_off[C] = overflow ? 0 : excl[(rank + 1) * C] - excl[rank * C];
    info[0] = overflow ? 0 : excl[(rank + 1) * C] - excl[rank * C];
    info[1] = s_max;
    info[2] = overflow ? 1 : 0;
  }
}

__global__ void set_segment_kernel(int64_t* seg_off, int64_t n) {
  seg_off[0] = 0;
  seg_off[1] = n;
}

}  // namespace

extern "C" {

uint32_t b2_wang_hash_u32(uint32_t key) { return wang_hash_u32(key); }

int b2_join_dest_rank(uint32_t key, int nranks) {
  if (nranks < 1 || (nranks & (nranks - 1)) != 0) return -1;  // the hash routes over a power of two
  if (nranks == 1) return 0;
  int bits = 0;
  while ((1 << bits) < nranks) ++bits;
  return (int)(wang_hash_u32(key) >> (32 - bits));
}

// Sizes cover the plan with and without the perfect-hash path (a ctx tunable decides between them).
size_t b2_join_ws_bytes(int64_t nl, int64_t nr) {
  if (nl < 0 || nr < 0) return 0;
  return std::max(make_plan(nl, nr, 0, 0).total, make_plan(nl, nr, 0, 0, false, 0).total);
}
size_t b2_join_ws_bytes_adjacent_outputs(int64_t nl, int64_t nr) {
  if (nl < 0 || nr < 0) return 0;
  return std::max(make_plan(nl, nr, 0, 0, true).total, make_plan(nl, nr, 0, 0, true, 0).total);
}
size_t b2_join_min_ws_bytes(int64_t nl, int64_t nr) {
  if (nl < 0 || nr < 0) return 0;
  return std::max(make_plan(nl, nr, 0, kMaxSliceBits).total, make_plan(nl, nr, 0, kMaxSliceBits, false, 0).total);
}

int b2_join_u32_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                    const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, uint32_t* d_out_fk,
                    uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                    uint64_t* d_out_rows, int hash_skip_bits, void* d_ws, size_t ws_bytes,
                    void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl == 0 || (d_fk && d_y), "null left column");
  B2_REQUIRE(ctx, nr == 0 || (d_pk && d_x), "null right column");
  PartInput lin, rin;
  lin.keys = d_fk;
  lin.vals = d_y;
  rin.keys = d_pk;
  rin.vals = d_x;
  return join_impl(ctx, lin, nl, rin, nr, d_out_fk, d_out_y, d_out_x, out_capacity, d_out_rows,
                   hash_skip_bits, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int b2_join_u32_phased_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                           const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, uint32_t* d_out_fk,
                           uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                           uint64_t* d_out_rows, int hash_skip_bits, int phases, void* d_ws, size_t ws_bytes,
                           void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl == 0 || (d_fk && d_y), "null left column");
  B2_REQUIRE(ctx, nr == 0 || (d_pk && d_x), "null right column");
  PartInput lin, rin;
  lin.keys = d_fk;
  lin.vals = d_y;
  rin.keys = d_pk;
  rin.vals = d_x;
  return join_impl(ctx, lin, nl, rin, nr, d_out_fk, d_out_y, d_out_x, out_capacity, d_out_rows,
                   hash_skip_bits, d_ws, ws_bytes, static_cast<cudaStream_t>(stream), nullptr, phases);
}

int b2_join_aggr_u32_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                         const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, int filter_y,
                         uint32_t y_threshold, b2_join_aggr* d_out, int hash_skip_bits, void* d_ws,
                         size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl == 0 || (d_fk && d_y), "null left column");
  B2_REQUIRE(ctx, nr == 0 || (d_pk && d_x), "null right column");
  PartInput lin, rin;
  lin.keys = d_fk;
  lin.vals = d_y;
  rin.keys = d_pk;
  rin.vals = d_x;
  JoinAggCfg agg;
  agg.y_thr = y_threshold;
  agg.filter_y = filter_y != 0;
  agg.d_out = d_out;
  return join_impl(ctx, lin, nl, rin, nr, nullptr, nullptr, nullptr, 0, nullptr, hash_skip_bits, d_ws,
                   ws_bytes, static_cast<cudaStream_t>(stream), &agg);
}

int b2_join_pairs_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, int64_t nl, const uint64_t* d_r_pairs,
                      int64_t nr, uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x,
                      int64_t out_capacity, uint64_t* d_out_rows, int hash_skip_bits, void* d_ws,
                      size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl == 0 || d_l_pairs, "null left pairs");
  B2_REQUIRE(ctx, nr == 0 || d_r_pairs, "null right pairs");
  PartInput lin, rin;
  lin.pairs = reinterpret_cast<const uint2*>(d_l_pairs);
  rin.pairs = reinterpret_cast<const uint2*>(d_r_pairs);
  // an empty side has no pairs pointer; part_full is never reached then (nl == 0 || nr == 0)
  return join_impl(ctx, lin, nl, rin, nr, d_out_fk, d_out_y, d_out_x, out_capacity, d_out_rows,
                   hash_skip_bits, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- multi-GPU shuffle: route rows to the GPU that owns their hash range --------------------
size_t b2_shuffle_ws_bytes(int64_t n, int nranks) {
  if (n < 0 || nranks < 1) return 0;
  int bits = 0;
  while ((1 << bits) < nranks) ++bits;
  return part_full_ws_bytes(n, bits) + 256;
}

int b2_shuffle_partition_u32_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n,
                                 int nranks, uint64_t* d_pairs_out, int64_t* d_dest_off, void* d_ws,
                                 size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0, "negative size");
  B2_REQUIRE(ctx, nranks >= 1 && nranks <= 1024 && (nranks & (nranks - 1)) == 0,
             "nranks must be a power of two <= 1024");
  B2_REQUIRE(ctx, d_dest_off != nullptr, "d_dest_off is null");
  B2_REQUIRE(ctx, n == 0 || (d_key && d_val && d_pairs_out), "null pointer");
  int bits = 0;
  while ((1 << bits) < nranks) ++bits;
  PartInput in;
  in.keys = d_key;
  in.vals = d_val;
  return part_full(ctx, in, n, bits, 0, 0, 0, 0, reinterpret_cast<uint2*>(d_pairs_out), nullptr, n,
                   d_dest_off, nullptr, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- fused multi-GPU shuffle: count, then scatter straight into the peers' receive buffers -----
size_t b2_shuffle_p2p_ws_bytes(int64_t n, int bits) {
  if (n < 0 || bits < 0 || bits > kPartMaxBits) return 0;
  return 256 + b2_align_up(part_pass_ws_bytes(n, 1, bits), 256);
}

int b2_shuffle_p2p_count_dev(b2_ctx* ctx, const uint32_t* d_key, int64_t n, int bits, int64_t* d_bucket_off,
                             void* d_ws, size_t ws_bytes, void* stream) {
  return b2_shuffle_p2p_count_lt_dev(ctx, d_key, nullptr, n, bits, 0, 0u, d_bucket_off, d_ws, ws_bytes, stream);
}

int b2_shuffle_p2p_count_lt_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n, int bits,
                                int filter_val, uint32_t val_threshold, int64_t* d_bucket_off, void* d_ws,
                                size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0 && bits >= 0 && bits <= kPartMaxBits, "bits must be in 0..10");
  B2_REQUIRE(ctx, d_bucket_off != nullptr && (n == 0 || d_key != nullptr), "null pointer");
  B2_REQUIRE(ctx, !filter_val || n == 0 || d_val != nullptr, "the predicate needs the value column");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
             "workspace must be 256 B aligned");
  if (ws_bytes < b2_shuffle_p2p_ws_bytes(n, bits))
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "p2p shuffle workspace", "use b2_shuffle_p2p_ws_bytes()");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(d_ws);
  int64_t* seg = reinterpret_cast<int64_t*>(base);
  set_segment_kernel<<<1, 1, 0, s>>>(seg, n);
  B2_LAUNCH_CHECK(ctx, "set_segment_kernel");
  PartInput in;
  in.keys = d_key;
  in.vals = filter_val ? d_val : nullptr;
  PartGeom g;
  g.bits = bits;
  g.val_pred = filter_val != 0;  // rows that fail `value < val_threshold` are not counted ...
  g.val_thr = val_threshold;
  return part_count(ctx, in, n, seg, 1, g, d_bucket_off, base + 256, ws_bytes - 256, s);
}

int b2_shuffle_p2p_plan_dev(b2_ctx* ctx, const int64_t* const* d_off_ptrs, const uint64_t* d_recv_base, int rank,
                            int nranks, int bits, int64_t capacity_rows, uint64_t* d_bucket_addr,
                            int64_t* d_seg_off, int64_t* d_info, const int64_t* d_prev_abort, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, bits >= 0 && bits <= kPartMaxBits, "bits must be in 0..10");
  B2_REQUIRE(ctx, nranks >= 1 && (nranks & (nranks - 1)) == 0 && nranks <= (1 << bits),
             "nranks must be a power of two <= 2^bits");
  B2_REQUIRE(ctx, rank >= 0 && rank < nranks, "rank out of range");
  B2_REQUIRE(ctx, d_off_ptrs && d_recv_base && d_bucket_addr && d_seg_off && d_info, "null pointer");
  shuffle_plan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      d_off_ptrs, d_recv_base, rank, nranks, bits, capacity_rows, d_bucket_addr, d_seg_off, d_info, d_prev_abort);
  B2_LAUNCH_CHECK(ctx, "shuffle_plan_kernel");
  return B2_OK;
}

int b2_shuffle_p2p_scatter_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n,
                               int bits, const uint64_t* d_bucket_addr, const int64_t* d_abort, void* d_ws,
                               size_t ws_bytes, void* stream) {
  return b2_shuffle_p2p_scatter_lt_dev(ctx, d_key, d_val, n, bits, 0, 0u, d_bucket_addr, d_abort, d_ws, ws_bytes,
                                       stream);
}

int b2_shuffle_p2p_scatter_lt_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n, int bits,
                                  int filter_val, uint32_t val_threshold, const uint64_t* d_bucket_addr,
                                  const int64_t* d_abort, void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0 && bits >= 0 && bits <= kPartMaxBits, "bits must be in 0..10");
  B2_REQUIRE(ctx, d_bucket_addr != nullptr && (n == 0 || (d_key && d_val)), "null pointer");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
             "workspace must be 256 B aligned");
  if (ws_bytes < b2_shuffle_p2p_ws_bytes(n, bits))
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "p2p shuffle workspace", "use b2_shuffle_p2p_ws_bytes()");
  char* base = static_cast<char*>(d_ws);
  PartInput in;
  in.keys = d_key;
  in.vals = d_val;
  PartGeom g;
  g.bits = bits;
  g.val_pred = filter_val != 0;  // ... and never cross the link
  g.val_thr = val_threshold;
  return part_scatter(ctx, in, n, reinterpret_cast<const int64_t*>(base), 1, g, nullptr, 0, d_bucket_addr,
                      nullptr, base + 256, ws_bytes - 256, static_cast<cudaStream_t>(stream), d_abort);
}

size_t b2_join_seg_ws_bytes(int64_t nl, int64_t nr, int hash_skip_bits, int seg_bits) {
  SegPlan P, Q;  // enough for the plan with and without the perfect-hash path (a ctx tunable decides)
  if (nl < 0 || nr < 0 || !make_seg_plan(nl, nr, hash_skip_bits, seg_bits, &P) ||
      !make_seg_plan(nl, nr, hash_skip_bits, seg_bits, &Q, 0, false))
    return 0;
  return std::max(P.total, Q.total);
}
size_t b2_join_seg_cap_ws_bytes(int64_t nl_cap, int64_t nr_cap, int64_t nr_expected, int hash_skip_bits,
                                int seg_bits) {
  SegPlan P, Q;
  if (nl_cap < 0 || nr_cap < 0 || !make_seg_plan(nl_cap, nr_cap, hash_skip_bits, seg_bits, &P, nr_expected) ||
      !make_seg_plan(nl_cap, nr_cap, hash_skip_bits, seg_bits, &Q, nr_expected, false))
    return 0;
  return std::max(P.total, Q.total);
}

int b2_join_pairs_seg_cap_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off, int64_t nl_cap,
                              const uint64_t* d_r_pairs, const int64_t* d_r_seg_off, int64_t nr_cap,
                              int64_t nr_expected, int seg_bits, uint32_t* d_out_fk, uint32_t* d_out_y,
                              uint32_t* d_out_x, int64_t out_capacity, uint64_t* d_out_rows, int hash_skip_bits,
                              const int64_t* d_abort, void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl_cap == 0 || d_l_pairs, "null left pairs");
  B2_REQUIRE(ctx, nr_cap == 0 || d_r_pairs, "null right pairs");
  B2_REQUIRE(ctx, out_capacity == 0 || (d_out_fk && d_out_y && d_out_x), "null output column");
  return join_seg_impl(ctx, reinterpret_cast<const uint2*>(d_l_pairs), d_l_seg_off, nl_cap,
                       reinterpret_cast<const uint2*>(d_r_pairs), d_r_seg_off, nr_cap, seg_bits, d_out_fk,
                       d_out_y, d_out_x, out_capacity, d_out_rows, hash_skip_bits, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream), nr_expected, d_abort);
}

int b2_join_pairs_seg_cap_phased_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off,
                                     int64_t nl_cap, const uint64_t* d_r_pairs, const int64_t* d_r_seg_off,
                                     int64_t nr_cap, int64_t nr_expected, int seg_bits, uint32_t* d_out_fk,
                                     uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                                     uint64_t* d_out_rows, int hash_skip_bits, const int64_t* d_abort, int phases,
                                     void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl_cap == 0 || d_l_pairs || !(phases & 2), "null left pairs");
  B2_REQUIRE(ctx, nr_cap == 0 || d_r_pairs, "null right pairs");
  B2_REQUIRE(ctx, out_capacity == 0 || (d_out_fk && d_out_y && d_out_x) || !(phases & 2), "null output column");
  return join_seg_impl(ctx, reinterpret_cast<const uint2*>(d_l_pairs), d_l_seg_off, nl_cap,
                       reinterpret_cast<const uint2*>(d_r_pairs), d_r_seg_off, nr_cap, seg_bits, d_out_fk,
                       d_out_y, d_out_x, out_capacity, d_out_rows, hash_skip_bits, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream), nr_expected, d_abort, phases);
}

int b2_join_aggr_pairs_seg_cap_phased_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off,
                                          int64_t nl_cap, const uint64_t* d_r_pairs, const int64_t* d_r_seg_off,
                                          int64_t nr_cap, int64_t nr_expected, int seg_bits, int filter_y,
                                          uint32_t y_threshold, b2_join_aggr* d_out, int hash_skip_bits,
                                          const int64_t* d_abort, int phases, void* d_ws, size_t ws_bytes,
                                          void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl_cap == 0 || d_l_pairs || !(phases & 2), "null left pairs");
  B2_REQUIRE(ctx, nr_cap == 0 || d_r_pairs, "null right pairs");
  B2_REQUIRE(ctx, d_out != nullptr, "null aggregate result");
  JoinAggCfg agg;
  agg.filter_y = filter_y != 0;
  agg.y_thr = y_threshold;
  agg.d_out = d_out;
  return join_seg_impl(ctx, reinterpret_cast<const uint2*>(d_l_pairs), d_l_seg_off, nl_cap,
                       reinterpret_cast<const uint2*>(d_r_pairs), d_r_seg_off, nr_cap, seg_bits, nullptr, nullptr,
                       nullptr, 0, nullptr, hash_skip_bits, d_ws, ws_bytes, static_cast<cudaStream_t>(stream),
                       nr_expected, d_abort, phases, &agg);
}

int b2_join_pairs_seg_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off, int64_t nl,
                          const uint64_t* d_r_pairs, const int64_t* d_r_seg_off, int64_t nr, int seg_bits,
                          uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                          uint64_t* d_out_rows, int hash_skip_bits, void* d_ws, size_t ws_bytes,
                          void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl == 0 || d_l_pairs, "null left pairs");
  B2_REQUIRE(ctx, nr == 0 || d_r_pairs, "null right pairs");
  B2_REQUIRE(ctx, out_capacity == 0 || (d_out_fk && d_out_y && d_out_x), "null output column");
  return join_seg_impl(ctx, reinterpret_cast<const uint2*>(d_l_pairs), d_l_seg_off, nl,
                       reinterpret_cast<const uint2*>(d_r_pairs), d_r_seg_off, nr, seg_bits, d_out_fk,
                       d_out_y, d_out_x, out_capacity, d_out_rows, hash_skip_bits, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

// ---- standalone partition (PartitionDpu) ----------------------------------------------------
size_t b2_partition_ws_bytes(int64_t n, int nparts) {
  if (n < 0 || nparts < 1) return 0;
  int bits = 0;
  while ((1 << bits) < nparts) ++bits;
  const size_t pairs = b2_align_up((size_t)n * 8, 256);
  return part_full_ws_bytes(n, bits) + pairs * (bits > kPartMaxBits ? 2 : 1) + 256;
}

int b2_partition_u32_dev(b2_ctx* ctx, const uint32_t* const* d_cols_in, uint32_t* const* d_cols_out,
                         int ncols, int64_t n, int nparts, int skip_bits, int64_t* d_part_off,
                         void* d_ws, size_t ws_bytes, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0 && n < (1ll << 32), "row count must fit the 32-bit row index");
  B2_REQUIRE(ctx, ncols >= 1 && ncols <= 16, "1..16 columns");
  B2_REQUIRE(ctx, nparts >= 1 && nparts <= (1 << (2 * kPartMaxBits)) && (nparts & (nparts - 1)) == 0,
             "nparts must be a power of two <= 2^20");
  B2_REQUIRE(ctx, skip_bits >= 0 && skip_bits <= 16, "skip_bits out of range");
  B2_REQUIRE(ctx, d_cols_in && d_cols_out && d_part_off, "null pointer");
  B2_REQUIRE(ctx, d_ws != nullptr && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
             "workspace must be 256 B aligned");
  if (ws_bytes < b2_partition_ws_bytes(n, nparts))
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "partition workspace", "use b2_partition_ws_bytes()");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bits = 0;
  while ((1 << bits) < nparts) ++bits;
  B2_REQUIRE(ctx, bits + skip_bits <= 32, "not enough hash bits");
  char* base = static_cast<char*>(d_ws);
  const size_t pairs_bytes = b2_align_up((size_t)n * 8, 256);
  uint2* pairs = reinterpret_cast<uint2*>(base);
  uint2* tmp = bits > kPartMaxBits ? reinterpret_cast<uint2*>(base + pairs_bytes) : nullptr;
  char* pws = base + pairs_bytes * (tmp ? 2 : 1);
  PartInput in;
  in.keys = d_cols_in[0];
  in.vals = nullptr;  // carry the row index: the permutation (selection_indices_vector)
  B2_RETURN_NOT_OK(part_full(ctx, in, n, bits, skip_bits, 0, 0, 0, pairs, tmp, n, d_part_off,
                             nullptr, pws, ws_bytes - (size_t)(pws - base), s));
  if (n > 0) {
    int64_t grid = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    for (int c = 0; c < ncols; ++c) {
      B2_REQUIRE(ctx, d_cols_in[c] && d_cols_out[c], "null column pointer");
      part_gather_kernel<<<(unsigned)grid, 256, 0, s>>>(pairs, n, d_cols_in[c], d_cols_out[c],
                                                        c == 0 ? 1 : 0);
      B2_LAUNCH_CHECK(ctx, "part_gather_kernel");
    }
  }
  return B2_OK;
}

}  // extern "C"
