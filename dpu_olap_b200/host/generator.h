// generator.h — the reference's synthetic inputs (host/generator/generator.cc:22-111 over the
// vendored copy of Arrow's testing/random.cc, :103-125,185-222,677-760), restated without
// libarrow_testing (absent from the image): the same seed stream
// (std::default_random_engine(42) -> uniform_int_distribution<int32_t>(1, INT32_MAX)), the same
// "bitmap first, data second" seed consumption, pcg32_fast and libstdc++'s
// uniform_int_distribution, so batches are bit-identical to the reference's.
#pragma once
#include <arrow/api.h>

#include <limits>
#include <random>
#include <string>
#include <vector>

namespace upmemeval {
namespace generator {

class RandomArrayGenerator {  // arrow::random::RandomArrayGenerator's seed stream
 public:
  using SeedType = int32_t;
  explicit RandomArrayGenerator(SeedType seed)
      : seed_distribution_(static_cast<SeedType>(1), std::numeric_limits<SeedType>::max()), seed_rng_(seed) {}
  SeedType seed() { return seed_distribution_(seed_rng_); }
  // Non-null uint32 array, values uniform in [min, max]
  std::shared_ptr<arrow::Array> UInt32(int64_t size, uint32_t min, uint32_t max);

 private:
  std::uniform_int_distribution<SeedType> seed_distribution_;
  std::default_random_engine seed_rng_;
};

// One array per field per batch, min/max from field metadata (random.cc:689-699,740-743)
arrow::RecordBatchVector MakeRandomRecordBatches(RandomArrayGenerator& g,
                                                 const std::shared_ptr<arrow::Schema>& schema,
                                                 int num_batches, int batch_size);
arrow::Result<arrow::ArrayVector> MakeIndexColumn(int num_batches, int batch_size);
arrow::RecordBatchVector AddColumn(const std::string& name, const arrow::RecordBatchVector& batches,
                                   arrow::ArrayVector column);  // inserted at index 0
arrow::Result<arrow::ArrayVector> MakeForeignKeyColumn(RandomArrayGenerator& g, uint32_t pk_batch_size,
                                                       int32_t num_batches, int32_t batch_size);

std::shared_ptr<arrow::Array> ArrayOf(const std::vector<uint32_t>& values);
std::shared_ptr<arrow::RecordBatch> RecordBatchOf(std::vector<std::string> names,
                                                  std::vector<std::shared_ptr<arrow::Array>> data);

}  // namespace generator
}  // namespace upmemeval
