// make_generator_golden.cc — produces tests/golden/generator_golden.json.
//
// Executes the reference generator's own statements (host/generator/random.cc:103-109 and
// arrow/testing/random.h's seed stream) with the REAL libstdc++ <random> and the REAL PCG header
// that Arrow vendors (arrow/vendored/pcg/pcg_random.hpp from the pyarrow wheel), so the oracle's
// restatement (oracle/olap_oracle.c) and the CUDA generator are pinned against what the
// reference would execute — not against another restatement.
//
// Build + run (see tests/golden/make_golden.sh):
//   g++ -O2 -std=c++17 -I$PYARROW/include make_generator_golden.cc -o /tmp/mkgold && /tmp/mkgold
#include <cstdint>
#include <cstdio>
#include <limits>
#include <random>
#include <vector>

#include "arrow/vendored/pcg/pcg_random.hpp"

using SeedType = int32_t;  // arrow/testing/random.h
using pcg32_fast = ::arrow_vendored::pcg32_fast;

struct Gen {  // arrow::random::RandomArrayGenerator's seed stream
  explicit Gen(SeedType seed)
      : seed_distribution_(static_cast<SeedType>(1), std::numeric_limits<SeedType>::max()),
        seed_rng_(seed) {}
  SeedType seed() { return seed_distribution_(seed_rng_); }
  std::uniform_int_distribution<SeedType> seed_distribution_;
  std::default_random_engine seed_rng_;
};

// GenerateOptions<uint32_t, std::uniform_int_distribution<uint32_t>>: bitmap first (seed_++),
// then data (seed_++)  — random.cc:103-125,190-196
static std::vector<uint32_t> numeric_u32(SeedType seed, uint32_t lo, uint32_t hi, size_t n) {
  SeedType seed_ = seed;
  {
    pcg32_fast rng(seed_++);  // GenerateBitmap consumes one seed; its draws do not matter here
    (void)rng;
  }
  pcg32_fast rng(seed_++);
  std::uniform_int_distribution<uint32_t> dist(lo, hi);
  std::vector<uint32_t> out(n);
  for (auto& v : out) v = dist(rng);
  return out;
}

static void dump(const char* name, SeedType seed, uint32_t lo, uint32_t hi, size_t n, bool last) {
  auto v = numeric_u32(seed, lo, hi, n);
  uint64_t sum = 0, lt = 0, xr = 0;
  for (size_t i = 0; i < n; ++i) {
    sum += v[i];
    lt += v[i] < (1u << 30);
    xr ^= (uint64_t)v[i] * (i + 1);
  }
  std::printf("  \"%s\": {\"seed\": %d, \"lo\": %u, \"hi\": %u, \"n\": %zu, \"first\": [", name, seed,
              lo, hi, n);
  for (int i = 0; i < 8; ++i) std::printf("%u%s", v[i], i < 7 ? ", " : "");
  std::printf("], \"last\": %u, \"sum\": %llu, \"count_lt_2p30\": %llu, \"xor_weighted\": %llu}%s\n",
              v[n - 1], (unsigned long long)sum, (unsigned long long)lt, (unsigned long long)xr,
              last ? "" : ",");
}

int main() {
  std::printf("{\n");
  {
    Gen g(42);
    std::printf("  \"seed_stream_42\": [");
    for (int i = 0; i < 16; ++i) std::printf("%d%s", g.seed(), i < 15 ? ", " : "");
    std::printf("],\n");
  }
  {
    Gen g(7);
    std::printf("  \"seed_stream_7\": [");
    for (int i = 0; i < 8; ++i) std::printf("%d%s", g.seed(), i < 7 ? ", " : "");
    std::printf("],\n");
  }
  Gen g(42);
  const SeedType s0 = g.seed(), s1 = g.seed(), s2 = g.seed();
  dump("full_range_batch0", s0, 0u, 0xffffffffu, 65536, false);
  dump("full_range_batch1", s1, 0u, 0xffffffffu, 65536, false);
  dump("take_index_64k", s2, 0u, 65535u, 8192, false);                 // take_test.cc:58-63
  dump("fk_2p21_batch3", s2, 3u << 21, (4u << 21) - 1, 4096, false);   // generator.cc:50-54
  dump("pow2_span_1", s1, 5u, 5u, 16, false);
  dump("non_pow2_span_1000", s0, 10u, 1009u, 4096, false);             // rejection path
  dump("non_pow2_span_3e9", s1, 0u, 2999999999u, 4096, true);          // heavy rejection
  std::printf("}\n");
  return 0;
}
