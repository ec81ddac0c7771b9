// join_l2.cu — build and probe of the inner equi-join against a hash table that lives in the
// 126 MB L2 of the B200 instead of in shared memory.
//
// Replaces kernel_hash_build (dpu/shared/kernels/hash_build.c:9-35), kernel_hash_probe
// (hash_probe.c:9-46) and ht_put / ht_get (dpu/shared/hashtable/hashtable.c:89-192): the DPU keeps
// a 4 Mi-slot linear-probing table per partition in its 64 MB MRAM bank and pays 2-3 MRAM DMAs plus
// a hardware mutex per insert. Semantics are Arrow's inner hash join (join_native.cc:31-36), as in
// join.cu: unmatched probe rows are dropped, duplicate build keys emit one row per duplicate.
//
// Why L2: a shared-memory table holds ~4096 build rows, so 2^32 rows need 2^20 partitions = two
// radix passes over both sides (2 x 20 B per row and side) before a single row is joined. A table
// of 2^21 rows (~35 MB) stays resident in L2 while its group streams through, so ONE pass of at
// most 2^10 groups is enough up to 2^31 build rows — and small joins need no partitioning at all.
//
// Table layout: a bucket is one 32-byte sector, [count, pad, (key, value) x 3]. An insert is ONE
// atomicAdd on the count (the returned value is the slot; >= 3 means "full, go to the next
// bucket") plus one 8-byte store; no compare-and-swap, no empty-key marker, so all 2^32 keys are
// legal. A probe is ONE 256-bit load (LDG.E.256): count, three keys and their values; it moves on
// to the next bucket only when count > 3 (rows overflowed). Clearing the table = zeroing the
// counts.
//
// One persistent cooperative launch walks the groups: clear | grid sync | build | grid sync |
// probe | grid sync. Every phase splits its rows evenly over all CTAs. Build groups larger than
// the table's capacity (skew, heavy duplicates) are inserted in chunks, each chunk probed by the
// whole probe side of the group.
#include <cooperative_groups.h>

#include <algorithm>

#include "join_l2.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kT = 256;            // threads per CTA
constexpr int kW = kT / 32;
constexpr int kI = 8;              // rows per thread per round
constexpr int kRound = kT * kI;    // 2048 rows
constexpr int kCtasPerSm = 3;
constexpr int kSlots = kJoinL2SlotsPerBucket;

struct L2Args {
  PartInput rin, lin;
  const int64_t* roff;
  const int64_t* loff;
  int64_t ngroups;
  uint32_t* table;
  uint32_t nbuckets;
  int64_t cap_rows;
  uint32_t* out_fk;
  uint32_t* out_y;
  uint32_t* out_x;
  int64_t out_cap;
  JoinState* st;
};

struct Bucket {
  uint32_t w[8];  // count, pad, k0, v0, k1, v1, k2, v2
};

// The table is written (atomics, stores) and read in different phases of the SAME launch, so its
// loads must not be served from a stale L1 line: .cg caches in L2 only.
__device__ __forceinline__ Bucket ld_bucket(const uint32_t* p) {
  Bucket b;
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]),
                 "=r"(b.w[6]), "=r"(b.w[7])
               : "l"(p)
               : "memory");
  return b;
}

// Keys of one group share the top bits of wang_hash (that is what grouped them), so the table is
// indexed by an independent multiplicative hash, range-reduced to [0, nbuckets) by a multiply-high.
__device__ __forceinline__ uint32_t bucket_of(uint32_t key, uint32_t nbuckets) {
  return __umulhi(key * 0x9E3779B1u, nbuckets);
}

// Rows [row0 + q*kT + tid] for q < kI, zero beyond row1.
__device__ __forceinline__ void load_rows(const PartInput& in, int64_t row0, int64_t row1, uint32_t tid,
                                          uint32_t (&k)[kI], uint32_t (&v)[kI]) {
#pragma unroll
  for (int q = 0; q < kI; ++q) {
    const int64_t i = row0 + q * kT + tid;
    k[q] = 0;
    v[q] = 0;
    if (i < row1) {
      if (in.pairs) {
        const uint2 kv = ld_stream_v2(in.pairs + i);
        k[q] = kv.x;
        v[q] = kv.y;
      } else {
        k[q] = ld_stream_u32(in.keys + i);
        v[q] = ld_stream_u32(in.vals + i);
      }
    }
  }
}

// This CTA's share of rows [r0, r1): equal contiguous slices, 128-row aligned.
__device__ __forceinline__ void cta_slice(int64_t r0, int64_t r1, int64_t& s0, int64_t& s1) {
  const int64_t n = r1 - r0;
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + 127) & ~(int64_t)127;
  s0 = r0 + min(n, (int64_t)blockIdx.x * per);
  s1 = r0 + min(n, ((int64_t)blockIdx.x + 1) * per);
}

__device__ __forceinline__ void clear_phase(const L2Args& a) {
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t b = (int64_t)blockIdx.x * kT + threadIdx.x; b < (int64_t)a.nbuckets; b += stride)
    a.table[8 * b] = 0;
}

__device__ __forceinline__ void build_phase(const L2Args& a, int64_t c0, int64_t c1) {
  const uint32_t tid = threadIdx.x;
  int64_t s0, s1;
  cta_slice(c0, c1, s0, s1);
  for (int64_t t0 = s0; t0 < s1; t0 += kRound) {
    uint32_t k[kI], v[kI], b[kI], slot[kI];
    load_rows(a.rin, t0, s1, tid, k, v);
    // all first-choice counters of the round are bumped before any result is looked at
#pragma unroll
    for (int q = 0; q < kI; ++q) {
      b[q] = bucket_of(k[q], a.nbuckets);
      slot[q] = 0;
      if (t0 + q * kT + tid < s1) slot[q] = atomicAdd(&a.table[8 * (size_t)b[q]], 1u);
    }
    // rows that found their bucket full move on to the next one — all of them at once, so the
    // atomics of one step are in flight together instead of one dependent chain per row
    uint32_t pend = 0;
#pragma unroll
    for (int q = 0; q < kI; ++q) pend |= (slot[q] >= (uint32_t)kSlots ? 1u : 0u) << q;
    while (pend) {
#pragma unroll
      for (int q = 0; q < kI; ++q) {
        if (pend >> q & 1) {
          b[q] = b[q] + 1 == a.nbuckets ? 0 : b[q] + 1;
          slot[q] = atomicAdd(&a.table[8 * (size_t)b[q]], 1u);
        }
      }
#pragma unroll
      for (int q = 0; q < kI; ++q)
        if ((pend >> q & 1) && slot[q] < (uint32_t)kSlots) pend &= ~(1u << q);
    }
#pragma unroll
    for (int q = 0; q < kI; ++q) {
      if (t0 + q * kT + tid < s1)
        *reinterpret_cast<uint2*>(a.table + 8 * (size_t)b[q] + 2 + 2 * slot[q]) = make_uint2(k[q], v[q]);
    }
  }
}

// Matches of `key` among the valid slots of a bucket: bit i set <=> slot i holds the key.
__device__ __forceinline__ uint32_t hit_mask(const Bucket& c, uint32_t key) {
  const uint32_t n = c.w[0];
  return ((n > 0 && c.w[2] == key) ? 1u : 0u) | ((n > 1 && c.w[4] == key) ? 2u : 0u) |
         ((n > 2 && c.w[6] == key) ? 4u : 0u);
}
__device__ __forceinline__ uint32_t slot_value(const Bucket& c, int i) {
  return i == 0 ? c.w[3] : (i == 1 ? c.w[5] : c.w[7]);
}

struct ProbeSmem {
  uint32_t warp_cnt[kW];
  uint32_t warp_off[kW];
  unsigned long long base;
};

__device__ __forceinline__ void probe_phase(const L2Args& a, int64_t l0, int64_t l1, ProbeSmem& sm) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = lanemask_lt();
  int64_t s0, s1;
  cta_slice(l0, l1, s0, s1);
  for (int64_t t0 = s0; t0 < s1; t0 += kRound) {
    uint32_t k[kI], y[kI], x0[kI], m[kI];
    load_rows(a.lin, t0, s1, tid, k, y);
    // ---- look up: four independent 256-bit table loads in flight per thread ----
#pragma unroll
    for (int h = 0; h < kI; h += 4) {
      Bucket c[4];
      uint32_t b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        b[j] = bucket_of(k[h + j], a.nbuckets);
        c[j] = ld_bucket(a.table + 8 * (size_t)b[j]);
      }
      uint32_t pend = 0;  // rows whose bucket overflowed: their chain continues in the next bucket
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = h + j;
        m[q] = 0;
        x0[q] = 0;
        if (t0 + q * kT + tid < s1) {
          const uint32_t hit = hit_mask(c[j], k[q]);
          if (hit) {
            x0[q] = slot_value(c[j], __ffs(hit) - 1);
            m[q] = __popc(hit);
          }
          pend |= (c[j].w[0] > (uint32_t)kSlots ? 1u : 0u) << j;
        }
      }
      while (pend) {  // all continuing rows step together: their loads overlap
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (pend >> j & 1) {
            b[j] = b[j] + 1 == a.nbuckets ? 0 : b[j] + 1;
            c[j] = ld_bucket(a.table + 8 * (size_t)b[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = h + j;
          if (pend >> j & 1) {
            const uint32_t hit = hit_mask(c[j], k[q]);
            if (hit) {
              if (m[q] == 0) x0[q] = slot_value(c[j], __ffs(hit) - 1);
              m[q] += __popc(hit);
            }
            if (c[j].w[0] <= (uint32_t)kSlots) pend &= ~(1u << j);  // nothing overflowed: chain ends
          }
        }
      }
    }
    // ---- output positions in (warp, item, lane) order: one 64-bit atomic per CTA and round ----
    uint32_t wtotal = 0;
#pragma unroll
    for (int q = 0; q < kI; ++q) {
      const uint32_t multi = __ballot_sync(0xffffffffu, m[q] > 1);
      wtotal += multi ? __reduce_add_sync(0xffffffffu, m[q]) : __popc(__ballot_sync(0xffffffffu, m[q] == 1));
    }
    if (lane == 0) sm.warp_cnt[warp] = wtotal;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = lane < kW ? sm.warp_cnt[lane] : 0;
      uint32_t incl = w;
#pragma unroll
      for (int o = 1; o < kW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane < kW) sm.warp_off[lane] = incl - w;
      if (lane == kW - 1 && incl > 0) sm.base = atomicAdd(&a.st->out_rows, (unsigned long long)incl);
    }
    __syncthreads();
    unsigned long long pos = sm.base + sm.warp_off[warp];
#pragma unroll
    for (int q = 0; q < kI; ++q) {
      const uint32_t multi = __ballot_sync(0xffffffffu, m[q] > 1);
      if (multi == 0) {  // unique build keys: ranks come from one ballot
        const uint32_t one = __ballot_sync(0xffffffffu, m[q] == 1);
        const unsigned long long p = pos + __popc(one & lt);
        if (m[q] == 1 && (int64_t)p < a.out_cap) {
          st_stream_u32(a.out_fk + p, k[q]);
          st_stream_u32(a.out_y + p, y[q]);
          st_stream_u32(a.out_x + p, x0[q]);
        }
        pos += __popc(one);
      } else {  // duplicate build keys somewhere in this warp: enumerate every match
        uint32_t incl = m[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        unsigned long long p = pos + incl - m[q];
        pos += __shfl_sync(0xffffffffu, incl, 31);
        if (m[q] == 1) {
          if ((int64_t)p < a.out_cap) {
            a.out_fk[p] = k[q];
            a.out_y[p] = y[q];
            a.out_x[p] = x0[q];
          }
        } else if (m[q] > 1) {
          uint32_t bb = bucket_of(k[q], a.nbuckets);
          while (true) {
            const Bucket cur = ld_bucket(a.table + 8 * (size_t)bb);
            uint32_t hit = hit_mask(cur, k[q]);
            while (hit) {
              const int i = __ffs(hit) - 1;
              hit &= hit - 1;
              if ((int64_t)p < a.out_cap) {
                a.out_fk[p] = k[q];
                a.out_y[p] = y[q];
                a.out_x[p] = slot_value(cur, i);
              }
              ++p;
            }
            if (cur.w[0] <= (uint32_t)kSlots) break;
            bb = bb + 1 == a.nbuckets ? 0 : bb + 1;
          }
        }
      }
    }
    // warp_cnt / warp_off / base are rewritten only after the next round's first barrier
  }
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kT, kCtasPerSm) join_l2_kernel(const L2Args a) {
  __shared__ ProbeSmem sm;
  cg::grid_group grid = cg::this_grid();
  // phase clock of CTA 0 (diagnostics: where a join step spends its time)
  const bool clock = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t_prev = clock ? global_ns() : 0, acc[3] = {0, 0, 0}, nsync = 0;
  auto lap = [&](int phase) {
    if (clock) {
      const unsigned long long t = global_ns();
      acc[phase] += t - t_prev;
      t_prev = t;
      ++nsync;
    }
  };
  for (int64_t g = 0; g < a.ngroups; ++g) {
    const int64_t r0 = a.roff[g], r1 = a.roff[g + 1];
    const int64_t l0 = a.loff[g], l1 = a.loff[g + 1];
    if (r1 == r0 || l1 == l0) continue;  // inner join: nothing to emit (uniform over the grid)
    for (int64_t c0 = r0; c0 < r1; c0 += a.cap_rows) {
      clear_phase(a);
      grid.sync();
      lap(0);
      build_phase(a, c0, min(r1, c0 + a.cap_rows));
      grid.sync();
      lap(1);
      probe_phase(a, l0, l1, sm);
      grid.sync();  // every probe is done before the counts are zeroed again
      lap(2);
    }
  }
  if (clock) {
    a.st->phase_ns[0] += acc[0];
    a.st->phase_ns[1] += acc[1];
    a.st->phase_ns[2] += acc[2];
    a.st->phase_ns[3] += nsync;
  }
}

}  // namespace

JoinL2Geom join_l2_geom(int64_t group_rows, int fill_x100) {
  JoinL2Geom G;
  if (group_rows < 1) group_rows = 1;
  if (fill_x100 < 25) fill_x100 = 25;
  if (fill_x100 > 250) fill_x100 = 250;
  int64_t nb = (group_rows * 100 + fill_x100 - 1) / fill_x100;
  nb = std::max<int64_t>(nb, 256);
  nb = std::min<int64_t>(nb, (int64_t)1 << 31);
  G.nbuckets = (uint32_t)nb;
  // a chunk fills at most 90 % of the slots, so overflow chains stay short
  G.cap_rows = nb * kSlots * 9 / 10;
  G.table_bytes = (size_t)nb * 32;
  return G;
}

int join_l2_run(b2_ctx* ctx, const PartInput& rin, const int64_t* d_roff, const PartInput& lin,
                const int64_t* d_loff, int64_t ngroups, const JoinL2Geom& geom, void* d_table,
                uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_cap, JoinState* st,
                cudaStream_t s) {
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_table) & 31) == 0, "table must be 32 B aligned");
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    B2_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, join_l2_kernel, kT, 0));
    if (ctas_per_sm < 1)
      return b2_set_error(ctx, B2_ERR_CUDA, "join_l2_kernel", "kernel does not fit on an SM");
    ctas_per_sm = std::min(ctas_per_sm, kCtasPerSm);
  }
  L2Args a;
  a.rin = rin;
  a.lin = lin;
  a.roff = d_roff;
  a.loff = d_loff;
  a.ngroups = ngroups;
  a.table = static_cast<uint32_t*>(d_table);
  a.nbuckets = geom.nbuckets;
  a.cap_rows = geom.cap_rows;
  a.out_fk = d_out_fk;
  a.out_y = d_out_y;
  a.out_x = d_out_x;
  a.out_cap = out_cap;
  a.st = st;
  void* params[] = {&a};
  B2_CUDA_OK(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(join_l2_kernel),
                                              dim3((unsigned)(ctx->sm_count * ctas_per_sm)), dim3(kT),
                                              params, 0, s));
  B2_LAUNCH_CHECK(ctx, "join_l2_kernel");
  return B2_OK;
}
