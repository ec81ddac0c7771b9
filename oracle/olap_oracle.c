/*
 * olap_oracle.c — CPU restatement of dpu_olap's columnar operator path. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker the CUDA path is compared against. Nothing under dpu_olap_b200/ may
 * import, link or call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Where the arithmetic comes from:
 *   - The operators' semantics are those of the reference's "Native" classes, which delegate to
 *     Apache Arrow C++ (pinned apache-arrow-8.0.0 by /root/reference/Dependencies.cmake:47-56;
 *     the source is NOT vendored in the reference tree). Call sites restated here:
 *       filter  host/filter/filter_native.cc:52-66   less(field_ref("v"), literal(1<<30))
 *       sum     host/aggr/aggr_native.cc:68-73       aggregate "sum" over uint32 -> uint64
 *       take    host/take/take_native.cc:24-31       cp::Take(values, indices, NoBoundsCheck) per batch
 *       join    host/join/join_native.cc:31-40,75    INNER hashjoin fk = pk, drop pk
 *     Parity is pinned two ways (tests/test_oracle.py): against the reference's own known-answer
 *     tests (filter_test.cc:24-61, aggr_test.cc:24-35, take_test.cc:24-46, join_test.cc:40-80,
 *     partition_test.cc:21-57) and against Arrow Acero 24.0.0 (the pyarrow wheel in the image)
 *     running the same plans on generator(42) inputs (tests/golden/arrow_golden.json).
 *   - The generator restates host/generator/random.cc:103-109 (pcg32_fast + libstdc++
 *     uniform_int_distribution) and arrow/testing/random.h's seed stream; pinned against values
 *     produced by the real libstdc++ / PCG headers (tests/golden/generator_golden.json).
 *   - The partition hash is dpu/shared/kernels/partition.c:20-28,45-46.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Generator
 * ------------------------------------------------------------------------------------------ */

/* std::default_random_engine == std::minstd_rand0: x <- 16807 x mod (2^31 - 1), min 1, max m-1.
 * arrow/testing/random.h: `std::default_random_engine seed_rng_(seed)`; a seed of 0 becomes 1. */
typedef struct { uint64_t x; } orc_minstd;
void orc_minstd_seed(orc_minstd* g, uint32_t seed) {
  uint64_t s = seed % 2147483647u;
  g->x = s == 0 ? 1 : s;
}
static uint64_t minstd_next(orc_minstd* g) {
  g->x = (g->x * 16807ull) % 2147483647ull;
  return g->x;
}

/* libstdc++ (GCC >= 11) std::uniform_int_distribution<int32_t>(a, b)(minstd_rand0), restated from
 * bits/uniform_int_dist.h operator(): urng range = 2^31 - 3. uctype is 64-bit. */
static uint64_t minstd_uniform(orc_minstd* g, uint64_t urange /* b - a */) {
  const uint64_t urngmin = 1, urngrange = 2147483646ull - 1ull;
  uint64_t ret;
  if (urngrange > urange) { /* downscaling, generic fallback (URNG range is not 2^32-1 / 2^64-1) */
    const uint64_t uerange = urange + 1;
    const uint64_t scaling = urngrange / uerange;
    const uint64_t past = uerange * scaling;
    do ret = minstd_next(g) - urngmin; while (ret >= past);
    ret /= scaling;
  } else if (urngrange < urange) { /* upscaling */
    uint64_t tmp;
    do {
      const uint64_t uerngrange = urngrange + 1;
      tmp = uerngrange * minstd_uniform(g, urange / uerngrange);
      ret = tmp + (minstd_next(g) - urngmin);
    } while (ret > urange || ret < tmp);
  } else {
    ret = minstd_next(g) - urngmin;
  }
  return ret;
}

/* RandomArrayGenerator::seed(): uniform_int_distribution<int32_t>(1, INT32_MAX)(seed_rng_)
 * (arrow/testing/random.h, identical in 8.0.0 and 24.0.0). */
int32_t orc_next_seed(orc_minstd* g) {
  return (int32_t)(minstd_uniform(g, 2147483647ull - 1ull) + 1ull);
}

/* pcg32_fast = pcg_engines::mcg_xsh_rs_64_32 (arrow/vendored/pcg): state = seed | 3;
 * output uses the OLD state: ((s >> 22) ^ s) >> (22 + (s >> 61)); s *= 6364136223846793005. */
typedef struct { uint64_t s; } orc_pcg;
static void pcg_seed(orc_pcg* g, uint64_t seed) { g->s = seed | 3ull; }
static uint32_t pcg_next(orc_pcg* g) {
  const uint64_t s = g->s;
  g->s = s * 6364136223846793005ull;
  return (uint32_t)(((s >> 22) ^ s) >> (22 + (s >> 61)));
}

/* std::uniform_int_distribution<uint32_t>(lo, hi) over a 32-bit URNG: libstdc++ takes Lemire's
 * nearly-divisionless path (_S_nd<uint64_t>) when the span is smaller than 2^32, else the raw
 * draw. */
static uint32_t pcg_uniform(orc_pcg* g, uint32_t lo, uint32_t hi) {
  const uint32_t urange = hi - lo;
  if (urange == 0xffffffffu) return pcg_next(g) + lo;
  const uint32_t range = urange + 1;
  uint64_t product = (uint64_t)pcg_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (uint32_t)(-range) % range;
    while (low < threshold) {
      product = (uint64_t)pcg_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32) + lo;
}

/* GenerateTypedDataNoNan (random.cc:103-109) for one uint32 array. data_seed is the value fed to
 * pcg32_fast: the array's seed() + 1, because GenerateBitmap ran first with seed_++
 * (random.cc:111-125,190-196). The int32 seed is converted to the 64-bit state type. */
void orc_gen_u32(int64_t data_seed, uint32_t lo, uint32_t hi, int64_t n, uint32_t* out) {
  orc_pcg g;
  pcg_seed(&g, (uint64_t)data_seed);
  for (int64_t i = 0; i < n; ++i) out[i] = pcg_uniform(&g, lo, hi);
}

/* generator::MakeIndexColumn (generator.cc:59-71): 0,1,2,... across batches, uint32 wrap. */
void orc_iota_u32(uint64_t start, int64_t n, uint32_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = (uint32_t)(start + (uint64_t)i);
}

/* ---------------------------------------------------------------------------------------------
 * Operators
 * ------------------------------------------------------------------------------------------ */

/* Filter: keep v < thr, input order preserved (filter_native.cc:59; DPU filter.c:25). */
int64_t orc_filter_lt_u32(const uint32_t* in, int64_t n, uint32_t thr, uint32_t* out) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (in[i] < thr) out[m++] = in[i];
  return m;
}

/* Sum: uint32 -> uint64 (aggr_native.cc:68-73; DPU aggr/main.c:44-51). */
uint64_t orc_sum_u32(const uint32_t* in, int64_t n) {
  uint64_t s = 0;
  for (int64_t i = 0; i < n; ++i) s += in[i];
  return s;
}

/* Take: out[j] = values[indices[j]] within one batch, no bounds check (take_native.cc:27). */
void orc_take_u32(const uint32_t* values, const uint32_t* indices, int64_t n_idx, uint32_t* out) {
  for (int64_t j = 0; j < n_idx; ++j) out[j] = values[indices[j]];
}

/* Partition hash (partition.c:20-28) and radix bucket (partition.c:45-46). */
uint32_t orc_wang_hash_u32(uint32_t key) {
  key += ~(key << 15);
  key ^= (key >> 10);
  key += (key << 3);
  key ^= (key >> 6);
  key += ~(key << 11);
  key ^= (key >> 16);
  return key;
}
/* nparts must be a power of two >= 1; skip_bits top hash bits are discarded first. */
uint32_t orc_bucket(uint32_t key, uint32_t nparts, int skip_bits) {
  int bits = 0;
  while ((1u << bits) < nparts) ++bits;
  if (bits == 0) return 0;
  return (uint32_t)((orc_wang_hash_u32(key) << skip_bits) >> (32 - bits));
}
void orc_partition_ids(const uint32_t* keys, int64_t n, uint32_t nparts, int skip_bits, uint32_t* ids) {
  for (int64_t i = 0; i < n; ++i) ids[i] = orc_bucket(keys[i], nparts, skip_bits);
}

/* Inner hash join L.fk = R.pk, Arrow semantics (join_native.cc:31-36): every matching (l, r)
 * pair yields a row (fk, y, x); duplicates on the build side multiply; unmatched rows vanish.
 * Chained hash table over R; rows are emitted in probe (L) order, matches of one probe row in
 * reverse build order — the caller compares sorted multisets, as join_test.cc:27-38,76-77 does.
 * Returns the number of matching pairs; writes at most cap rows. */
int64_t orc_join_u32(const uint32_t* fk, const uint32_t* y, int64_t nl, const uint32_t* pk,
                     const uint32_t* x, int64_t nr, uint32_t* out_fk, uint32_t* out_y,
                     uint32_t* out_x, int64_t cap) {
  if (nl == 0 || nr == 0) return 0;
  uint64_t nb = 1;
  while (nb < (uint64_t)nr * 2) nb <<= 1;
  int64_t* head = (int64_t*)malloc(nb * sizeof(int64_t));
  int64_t* next = (int64_t*)malloc((size_t)nr * sizeof(int64_t));
  if (!head || !next) { free(head); free(next); return -1; }
  memset(head, 0xff, nb * sizeof(int64_t));
  for (int64_t r = 0; r < nr; ++r) {
    const uint64_t b = (uint64_t)(pk[r] * 2654435761u) & (nb - 1);
    next[r] = head[b];
    head[b] = r;
  }
  int64_t m = 0;
  for (int64_t l = 0; l < nl; ++l) {
    const uint64_t b = (uint64_t)(fk[l] * 2654435761u) & (nb - 1);
    for (int64_t r = head[b]; r >= 0; r = next[r]) {
      if (pk[r] == fk[l]) {
        if (m < cap) { out_fk[m] = fk[l]; out_y[m] = y[l]; out_x[m] = x[r]; }
        ++m;
      }
    }
  }
  free(head);
  free(next);
  return m;
}

/* Order-independent checksum of a (fk, y, x) multiset, for full-size parity where sorting both
 * sides is too slow: sum over rows of a 64-bit mix of the triple (mod 2^64). */
static uint64_t mix64(uint64_t z) {
  z ^= z >> 33; z *= 0xff51afd7ed558ccdull;
  z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ull;
  z ^= z >> 33;
  return z;
}
uint64_t orc_triple_checksum(const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t n) {
  uint64_t s = 0;
  for (int64_t i = 0; i < n; ++i)
    s += mix64(((uint64_t)a[i] << 32 | b[i]) ^ mix64((uint64_t)c[i] + 0x9e3779b97f4a7c15ull));
  return s;
}
