#include "generator.h"

#include <arrow/util/key_value_metadata.h>
#include <arrow/vendored/pcg/pcg_random.hpp>

namespace upmemeval {
namespace generator {

std::shared_ptr<arrow::Array> RandomArrayGenerator::UInt32(int64_t size, uint32_t min, uint32_t max) {
  // GenerateOptions: the validity bitmap is drawn first and consumes one seed even when
  // null_probability is 0; the data generator uses the next one (random.cc:111-125,190-196)
  SeedType s = seed();
  s++;
  ::arrow_vendored::pcg32_fast rng(s++);
  std::uniform_int_distribution<uint32_t> dist(min, max);
  auto buf = arrow::AllocateBuffer(size * 4).ValueOrDie();
  uint32_t* data = reinterpret_cast<uint32_t*>(buf->mutable_data());
  for (int64_t i = 0; i < size; ++i) data[i] = dist(rng);
  return std::make_shared<arrow::UInt32Array>(size, std::shared_ptr<arrow::Buffer>(std::move(buf)));
}

arrow::RecordBatchVector MakeRandomRecordBatches(RandomArrayGenerator& g,
                                                 const std::shared_ptr<arrow::Schema>& schema,
                                                 int num_batches, int batch_size) {
  arrow::RecordBatchVector out;
  for (int b = 0; b < num_batches; ++b) {
    arrow::ArrayVector arrays;
    for (const auto& f : schema->fields()) {
      uint32_t lo = 0, hi = std::numeric_limits<uint32_t>::max();
      if (f->metadata()) {
        auto mn = f->metadata()->Get("min");
        auto mx = f->metadata()->Get("max");
        if (mn.ok()) lo = static_cast<uint32_t>(std::stoull(*mn));
        if (mx.ok()) hi = static_cast<uint32_t>(std::stoull(*mx));
      }
      arrays.push_back(g.UInt32(batch_size, lo, hi));
    }
    out.push_back(arrow::RecordBatch::Make(schema, batch_size, std::move(arrays)));
  }
  return out;
}

arrow::Result<arrow::ArrayVector> MakeIndexColumn(int num_batches, int batch_size) {
  arrow::ArrayVector out;
  uint32_t value = 0;  // a uint32 counter, as the reference's (generator.cc:60,66): wraps above 2^32
  for (int b = 0; b < num_batches; ++b) {
    ARROW_ASSIGN_OR_RAISE(auto buf, arrow::AllocateBuffer(static_cast<int64_t>(batch_size) * 4));
    uint32_t* data = reinterpret_cast<uint32_t*>(buf->mutable_data());
    for (int i = 0; i < batch_size; ++i) data[i] = value++;
    out.push_back(std::make_shared<arrow::UInt32Array>(batch_size, std::shared_ptr<arrow::Buffer>(std::move(buf))));
  }
  return out;
}

arrow::RecordBatchVector AddColumn(const std::string& name, const arrow::RecordBatchVector& batches,
                                   arrow::ArrayVector column) {
  arrow::RecordBatchVector out;
  for (size_t b = 0; b < batches.size(); ++b)
    out.push_back(batches[b]->AddColumn(0, arrow::field(name, arrow::uint32(), false), column[b]).ValueOrDie());
  return out;
}

arrow::Result<arrow::ArrayVector> MakeForeignKeyColumn(RandomArrayGenerator& g, uint32_t pk_batch_size,
                                                       int32_t num_batches, int32_t batch_size) {
  arrow::ArrayVector out;
  for (int32_t i = 0; i < num_batches; ++i)  // uniform in the matching pk batch's range (generator.cc:46-57)
    out.push_back(g.UInt32(batch_size, static_cast<uint32_t>(i) * pk_batch_size,
                           static_cast<uint32_t>(i + 1) * pk_batch_size - 1));
  return out;
}

std::shared_ptr<arrow::Array> ArrayOf(const std::vector<uint32_t>& values) {
  arrow::UInt32Builder b;
  if (!b.AppendValues(values).ok()) abort();
  return b.Finish().ValueOrDie();
}

std::shared_ptr<arrow::RecordBatch> RecordBatchOf(std::vector<std::string> names,
                                                  std::vector<std::shared_ptr<arrow::Array>> data) {
  arrow::FieldVector fields;
  for (size_t i = 0; i < names.size(); ++i) fields.push_back(arrow::field(names[i], data[i]->type(), false));
  const int64_t rows = data[0]->length();  // before the move below
  return arrow::RecordBatch::Make(arrow::schema(fields), rows, std::move(data));
}

}  // namespace generator
}  // namespace upmemeval
