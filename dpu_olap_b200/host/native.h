// native.h — the reference's "Native" operators (Arrow Acero on the host CPU), restated for the
// Arrow 24 C++ API that ships in this image (the reference pins Arrow 8.0.0 and does not compile
// against it — DESIGN.md §5). Same plans and options as host/filter/filter_native.cc:36-84,
// host/aggr/aggr_native.cc:39-93, host/take/take_native.cc:18-38, host/join/join_native.cc:14-80.
// They are the C++ CPU baseline and the differential oracle of host_test.cc.
#pragma once
#include <arrow/api.h>

#include <memory>

namespace upmemeval {

// Registers the compute kernels (Arrow >= 21 keeps them in libarrow_compute) and sizes the CPU
// thread pool (the benchmarks call arrow::SetCpuThreadPoolCapacity, filter_benchmark.cc:91).
arrow::Status InitNative(int threads);

namespace filter {
class FilterNative {
 public:
  FilterNative(std::shared_ptr<arrow::Schema> schema, arrow::RecordBatchVector batches)
      : schema_(std::move(schema)), batches_(std::move(batches)) {}
  arrow::Status Prepare();
  arrow::Result<std::shared_ptr<arrow::Table>> GetResult();
  arrow::Result<uint64_t> Run();

 private:
  std::shared_ptr<arrow::Schema> schema_;
  arrow::RecordBatchVector batches_;
  std::shared_ptr<arrow::Table> input_;
};
}  // namespace filter

namespace aggr {
class SumNative {  // AggrNative<arrow::UInt64Array> with the "sum" function (aggr_native.cc:68-70)
 public:
  SumNative(std::shared_ptr<arrow::Schema> schema, arrow::RecordBatchVector batches)
      : schema_(std::move(schema)), batches_(std::move(batches)) {}
  arrow::Status Prepare();
  arrow::Result<uint64_t> Run();

 private:
  std::shared_ptr<arrow::Schema> schema_;
  arrow::RecordBatchVector batches_;
  std::shared_ptr<arrow::Table> input_;
};
}  // namespace aggr

namespace take {
class TakeNative {
 public:
  TakeNative(std::shared_ptr<arrow::Schema> schema, arrow::RecordBatchVector batches,
             arrow::RecordBatchVector indices_batches)
      : schema_(std::move(schema)), batches_(std::move(batches)), indices_batches_(std::move(indices_batches)) {}
  arrow::Status Prepare() { return arrow::Status::OK(); }
  arrow::Result<std::shared_ptr<arrow::Table>> Run();

 private:
  std::shared_ptr<arrow::Schema> schema_;
  arrow::RecordBatchVector batches_, indices_batches_;
};
}  // namespace take

namespace join {
class JoinNative {
 public:
  JoinNative(std::shared_ptr<arrow::Schema> left_schema, std::shared_ptr<arrow::Schema> right_schema,
             arrow::RecordBatchVector left_batches, arrow::RecordBatchVector right_batches)
      : left_schema_(std::move(left_schema)),
        right_schema_(std::move(right_schema)),
        left_batches_(std::move(left_batches)),
        right_batches_(std::move(right_batches)) {}
  arrow::Status Prepare();
  arrow::Result<std::shared_ptr<arrow::Table>> Run();  // (fk, y, x): pk dropped (join_native.cc:75)

 private:
  std::shared_ptr<arrow::Schema> left_schema_, right_schema_;
  arrow::RecordBatchVector left_batches_, right_batches_;
  std::shared_ptr<arrow::Table> left_, right_;
};
}  // namespace join

}  // namespace upmemeval
