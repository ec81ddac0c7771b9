// operators.h — FilterGpu / SumGpu / TakeGpu / JoinGpu / PartitionGpu: siblings of the reference's
// *Dpu operator classes with the same constructor arguments and Prepare() / Run() / Timers()
// signatures (host/filter/filter_dpu.h:14-29, host/aggr/aggr_dpu.h:14-27,
// host/take/take_dpu.h:14-29, host/join/join_dpu.h:14-48, host/partition/partition_dpu.h:13-38),
// over libb200olap.so instead of dpu::DpuSet. Filter, Sum and Take also accept columns with nulls
// (Arrow's semantics, i.e. what the reference's *Native classes compute); join and partition keys
// must be non-null.
#pragma once
#include <arrow/api.h>

#include <memory>
#include <string>

#include "gpu_set.h"

namespace upmemeval {

namespace filter {
class FilterGpu {
 public:
  FilterGpu(gpu::GpuSet& system, arrow::RecordBatchVector batches, uint32_t threshold = 1u << 30)
      : system_(system), batches_(std::move(batches)), threshold_(threshold) {}
  // Columns of the other fixed-width types the library filters (int32 / float32, uint64 / int64 /
  // float64; the reference fixes T = uint32_t, dpu/shared/common.h:3): `v < threshold` in the column's
  // type, the threshold cast to it. Nullable columns welcome; the result keeps the column's type.
  FilterGpu(gpu::GpuSet& system, arrow::RecordBatchVector batches, std::shared_ptr<arrow::Scalar> threshold)
      : system_(system), batches_(std::move(batches)), threshold_(0), typed_threshold_(std::move(threshold)) {}
  arrow::Status Prepare();
  arrow::Result<std::shared_ptr<arrow::ChunkedArray>> GetResult();
  arrow::Result<uint64_t> Run();
  std::shared_ptr<timer::Timers> Timers() { return timers_; }

 private:
  gpu::GpuSet& system_;
  arrow::RecordBatchVector batches_;
  arrow::Result<std::shared_ptr<arrow::ChunkedArray>> GetTypedResult(const std::shared_ptr<arrow::DataType>& type);
  uint32_t threshold_;
  std::shared_ptr<arrow::Scalar> typed_threshold_;
  std::shared_ptr<timer::Timers> timers_;
};
}  // namespace filter

namespace aggr {
class SumGpu {
 public:
  SumGpu(gpu::GpuSet& system, arrow::RecordBatchVector batches)
      : system_(system), batches_(std::move(batches)) {}
  arrow::Status Prepare();
  arrow::Result<uint64_t> Run();
  // sum / count / min / max of the valid rows in one pass (nullable columns welcome; the reference's
  // enum AggregatorType only has AggrSum, shared/umq/kernels.h:22-25)
  arrow::Result<b2_aggr_u32> Aggregates();
  std::shared_ptr<timer::Timers> Timers() { return timers_; }

 private:
  gpu::GpuSet& system_;
  arrow::RecordBatchVector batches_;
  std::shared_ptr<timer::Timers> timers_;
};
}  // namespace aggr

namespace take {
class TakeGpu {
 public:
  TakeGpu(gpu::GpuSet& system, arrow::RecordBatchVector batches,
          arrow::RecordBatchVector indices_batches)
      : system_(system), batches_(std::move(batches)), indices_batches_(std::move(indices_batches)) {}
  arrow::Status Prepare();
  arrow::Result<std::shared_ptr<arrow::Table>> Run();
  std::shared_ptr<timer::Timers> Timers() { return timers_; }

 private:
  gpu::GpuSet& system_;
  arrow::RecordBatchVector batches_;
  arrow::RecordBatchVector indices_batches_;
  std::shared_ptr<timer::Timers> timers_;
};
}  // namespace take

namespace join {
class JoinGpu {
 public:
  JoinGpu(gpu::GpuSet& system, std::shared_ptr<arrow::Schema> left_schema,
          std::shared_ptr<arrow::Schema> right_schema, arrow::RecordBatchVector left_batches,
          arrow::RecordBatchVector right_batches)
      : system_(system),
        left_schema_(std::move(left_schema)),
        right_schema_(std::move(right_schema)),
        left_batches_(std::move(left_batches)),
        right_batches_(std::move(right_batches)) {}
  arrow::Status Prepare();
  // (fk, left payload, right payload); row order unspecified, as JoinDpu's (join_dpu.cc:376-399)
  arrow::Result<std::shared_ptr<arrow::Table>> Run();
  // fused [filter left payload < threshold ->] join -> COUNT / SUM / SUM, nothing materialised
  arrow::Result<b2_join_aggr> RunAggregate(bool filter_left_payload = false, uint32_t threshold = 0);
  std::shared_ptr<timer::Timers> Timers() { return timers_; }

 private:
  gpu::GpuSet& system_;
  std::shared_ptr<arrow::Schema> left_schema_, right_schema_;
  arrow::RecordBatchVector left_batches_, right_batches_;
  std::shared_ptr<timer::Timers> timers_;
};
}  // namespace join

namespace partition {
class PartitionGpu {
 public:
  PartitionGpu(gpu::GpuSet& system, std::shared_ptr<arrow::Schema> schema,
               arrow::RecordBatchVector batches, uint64_t nr_partitions,
               const std::string& partition_key)
      : system_(system),
        schema_(std::move(schema)),
        batches_(std::move(batches)),
        nr_partitions_(nr_partitions),
        partition_key_(partition_key) {}
  arrow::Status Prepare();
  arrow::Result<arrow::RecordBatchVector> Run();  // one record batch per partition
  std::shared_ptr<timer::Timers> Timers() { return timers_; }

 private:
  gpu::GpuSet& system_;
  std::shared_ptr<arrow::Schema> schema_;
  arrow::RecordBatchVector batches_;
  uint64_t nr_partitions_;
  std::string partition_key_;
  std::shared_ptr<timer::Timers> timers_;
};
}  // namespace partition

}  // namespace upmemeval
