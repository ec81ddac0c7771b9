// operators.cc — the *Gpu operator classes: Arrow record batches in, Arrow results out, one C-ABI
// call per leg. Column access follows the reference (data buffer #1 of a non-null fixed-width
// column, host/dpuext/arrow_utils.cc:23,60-66); result buffers are allocated by the caller with
// arrow::AllocateBuffer and filled by the library (filter_dpu.cc:34-39,79-83).
#include "operators.h"

#include <arrow/compute/api.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace upmemeval {
namespace {

arrow::Result<const uint32_t*> U32Values(const std::shared_ptr<arrow::Array>& col, bool allow_nulls) {
  if (col->type_id() != arrow::Type::UINT32)
    return arrow::Status::TypeError("expected a uint32 column, got ", col->type()->ToString());
  if (!allow_nulls && col->null_count() != 0)
    return arrow::Status::NotImplemented("join and partition columns must be non-null (the reference "
                                         "passes a nullptr validity bitmap, filter_dpu.cc:91)");
  return col->data()->GetValues<uint32_t>(1);
}

// Per-batch data pointers and lengths; with allow_nulls also the validity bitmaps (buffer #0) and
// their bit offsets (the array's offset), which the nullable entry points take.
struct ColumnPtrs {
  std::vector<const uint32_t*> ptrs;
  std::vector<int64_t> lens;
  std::vector<const uint8_t*> valid;
  std::vector<int64_t> valid_off;
  bool has_nulls = false;
  arrow::Status Append(const arrow::RecordBatchVector& batches, int column, bool allow_nulls = false) {
    for (const auto& b : batches) {
      if (column < 0 || column >= b->num_columns()) return arrow::Status::Invalid("no such column");
      const auto& col = b->column(column);
      ARROW_ASSIGN_OR_RAISE(const uint32_t* p, U32Values(col, allow_nulls));
      ptrs.push_back(p);
      lens.push_back(b->num_rows());
      const bool nulls = col->null_count() != 0;
      valid.push_back(nulls ? col->null_bitmap_data() : nullptr);
      valid_off.push_back(nulls ? col->offset() : 0);
      has_nulls = has_nulls || nulls;
    }
    return arrow::Status::OK();
  }
};

// Result array over page-locked memory from the GpuSet's recycling pool (a DMA target; pageable
// Arrow buffers would be filled through the driver's staging copies and page faults).
arrow::Result<std::shared_ptr<arrow::Array>> PinnedU32(gpu::GpuSet& sys, int64_t rows, uint32_t** data) {
  ARROW_ASSIGN_OR_RAISE(auto buf, sys.pinned().Acquire(std::max<int64_t>(rows, 1) * 4));
  *data = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(buf->data()));
  return std::make_shared<arrow::UInt32Array>(rows, arrow::SliceBuffer(buf, 0, rows * 4));
}

}  // namespace

// ---- Filter (FilterDpu, host/filter/filter_dpu.cc:23-174) ---------------------------------------
namespace filter {

arrow::Status FilterGpu::Prepare() {
  timers_ = std::make_shared<timer::Timers>();
  return arrow::Status::OK();
}

// int32 / float32 and 64-bit columns: one call on member 0 (b2_filter_lt_32_host_into /
// b2_filter_lt_64_host_into take the validity bitmaps too), result chunks sliced from one pinned slab.
arrow::Result<std::shared_ptr<arrow::ChunkedArray>> FilterGpu::GetTypedResult(
    const std::shared_ptr<arrow::DataType>& type) {
  b2_ctx* ctx = system_.ctx();
  int dtype = -1;
  switch (type->id()) {
    case arrow::Type::UINT32: dtype = B2_U32; break;
    case arrow::Type::INT32: dtype = B2_I32; break;
    case arrow::Type::FLOAT: dtype = B2_F32; break;
    case arrow::Type::UINT64: dtype = B2_U64; break;
    case arrow::Type::INT64: dtype = B2_I64; break;
    case arrow::Type::DOUBLE: dtype = B2_F64; break;
    default: return arrow::Status::TypeError("FilterGpu: no kernel for ", type->ToString(), " columns");
  }
  const int width = type->byte_width();
  // the threshold in the column's type, as raw bits
  std::shared_ptr<arrow::Scalar> thr = typed_threshold_ ? typed_threshold_ : arrow::MakeScalar(threshold_);
  if (!thr->type->Equals(*type)) {
    ARROW_ASSIGN_OR_RAISE(auto cast, arrow::compute::Cast(arrow::Datum(thr), type));
    thr = cast.scalar();
  }
  if (!thr->is_valid) return arrow::Status::Invalid("FilterGpu: null threshold");
  uint64_t bits = 0;
  const auto view = static_cast<const arrow::internal::PrimitiveScalarBase&>(*thr).view();
  if (static_cast<int>(view.size()) != width) return arrow::Status::Invalid("FilterGpu: threshold width");
  std::memcpy(&bits, view.data(), view.size());
  std::vector<const void*> ptrs;
  std::vector<const uint8_t*> valid;
  std::vector<int64_t> lens, valid_off;
  int64_t rows = 0;
  for (const auto& b : batches_) {
    const auto& col = b->column(0);
    if (!col->type()->Equals(*type)) return arrow::Status::TypeError("FilterGpu: batches of different types");
    const uint8_t* base = col->data()->buffers[1] ? col->data()->buffers[1]->data() : nullptr;
    ptrs.push_back(base ? base + col->offset() * width : nullptr);
    lens.push_back(b->num_rows());
    const bool nulls = col->null_count() != 0;
    valid.push_back(nulls ? col->null_bitmap_data() : nullptr);
    valid_off.push_back(nulls ? col->offset() : 0);
    rows += b->num_rows();
  }
  const int64_t nb = static_cast<int64_t>(ptrs.size());
  ARROW_ASSIGN_OR_RAISE(auto slab, system_.pinned().Acquire(std::max<int64_t>(rows, 1) * width));
  void* out = const_cast<uint8_t*>(slab->data());
  std::vector<int64_t> counts(nb > 0 ? nb : 1);
  uint64_t total = 0;
  b2_timings t{};
  if (width == 8) {
    B2_ARROW_RETURN_NOT_OK(ctx, b2_filter_lt_64_host_into(ctx, ptrs.data(), valid.data(), valid_off.data(), lens.data(),
                                                          nb, dtype, bits, out, rows, counts.data(), &total, &t));
  } else {
    B2_ARROW_RETURN_NOT_OK(ctx, b2_filter_lt_32_host_into(ctx, ptrs.data(), valid.data(), valid_off.data(), lens.data(),
                                                          nb, dtype, static_cast<uint32_t>(bits), out, rows,
                                                          counts.data(), &total, &t));
  }
  arrow::ArrayVector chunks;
  int64_t off = 0;
  for (int64_t b = 0; b < nb; ++b) {
    auto piece = arrow::SliceBuffer(slab, off * width, counts[b] * width);
    chunks.push_back(arrow::MakeArray(arrow::ArrayData::Make(type, counts[b], {nullptr, std::move(piece)}, 0)));
    off += counts[b];
  }
  timers_->Add(t);
  return arrow::ChunkedArray::Make(std::move(chunks), type);
}

arrow::Result<std::shared_ptr<arrow::ChunkedArray>> FilterGpu::GetResult() {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  if (!batches_.empty() && batches_[0]->num_columns() > 0) {
    const auto& type = batches_[0]->column(0)->type();
    if (typed_threshold_ || type->id() != arrow::Type::UINT32) return GetTypedResult(type);
  }
  ColumnPtrs in;
  ARROW_RETURN_NOT_OK(in.Append(batches_, 0, /*allow_nulls=*/true));
  const int64_t nb = static_cast<int64_t>(in.ptrs.size());
  int64_t rows = 0;
  for (int64_t l : in.lens) rows += l;
  // One page-locked slab receives the whole compacted result (streamed: upload, kernels and
  // download of successive groups of batches overlap); the chunks of the ChunkedArray are
  // zero-copy slices of it, one per input batch, in batch order (filter_dpu.cc:89-96,162-166).
  ARROW_ASSIGN_OR_RAISE(auto slab, system_.pinned().Acquire(std::max<int64_t>(rows, 1) * 4));
  std::vector<int64_t> counts(nb > 0 ? nb : 1);
  uint64_t total = 0;
  b2_timings t{};
  uint32_t* out = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(slab->data()));
  if (in.has_nulls) {  // Acero drops rows whose predicate is null: the result itself has no nulls
    B2_ARROW_RETURN_NOT_OK(ctx, b2_filter_lt_u32_nullable_host_into(
                                    ctx, in.ptrs.data(), in.valid.data(), in.valid_off.data(), in.lens.data(),
                                    nb, threshold_, out, rows, counts.data(), &total, &t));
  } else if (system_.size() > 1) {
    // several GPUs: every GPU filters its batch range and keeps the result; the chunks can only be
    // placed once all counts are known (the reference reads "output_buffer_length" first, then pulls
    // "output_buffer", filter_dpu.cc:57-83)
    b2_set* set = system_.set();
    B2_SET_RETURN_NOT_OK(set, b2_set_filter_lt_u32_host(set, in.ptrs.data(), in.lens.data(), nb, threshold_,
                                                        counts.data(), &total, &t));
    std::vector<uint32_t*> outs(nb > 0 ? nb : 1);
    int64_t o = 0;
    for (int64_t b = 0; b < nb; ++b) {
      outs[b] = out + o;
      o += counts[b];
    }
    b2_timings t2{};
    B2_SET_RETURN_NOT_OK(set, b2_set_filter_fetch_host(set, outs.data(), nb, &t2));
    timers_->Add(t2);
  } else {
    B2_ARROW_RETURN_NOT_OK(ctx, b2_filter_lt_u32_host_into(ctx, in.ptrs.data(), in.lens.data(), nb, threshold_,
                                                           out, rows, counts.data(), &total, &t));
  }
  arrow::ArrayVector chunks;
  int64_t off = 0;
  for (int64_t b = 0; b < nb; ++b) {
    auto piece = arrow::SliceBuffer(slab, off * 4, counts[b] * 4);
    chunks.push_back(std::make_shared<arrow::UInt32Array>(counts[b], std::move(piece)));
    off += counts[b];
  }
  timers_->Add(t);
  return arrow::ChunkedArray::Make(std::move(chunks), arrow::uint32());
}

arrow::Result<uint64_t> FilterGpu::Run() {
  ARROW_ASSIGN_OR_RAISE(auto result, GetResult());
  return static_cast<uint64_t>(result->length());
}

}  // namespace filter

// ---- Sum (SumDpu, host/aggr/aggr_dpu.cc:31-89) --------------------------------------------------
namespace aggr {

arrow::Status SumGpu::Prepare() {
  timers_ = std::make_shared<timer::Timers>();
  return arrow::Status::OK();
}

arrow::Result<b2_aggr_u32> SumGpu::Aggregates() {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  ColumnPtrs in;
  ARROW_RETURN_NOT_OK(in.Append(batches_, 0, /*allow_nulls=*/true));
  b2_aggr_u32 out{};
  b2_timings t{};
  B2_ARROW_RETURN_NOT_OK(ctx, b2_aggr_u32_host(ctx, in.ptrs.data(), in.valid.data(), in.valid_off.data(),
                                               in.lens.data(), static_cast<int64_t>(in.ptrs.size()), &out, &t));
  timers_->Add(t);
  return out;
}

arrow::Result<uint64_t> SumGpu::Run() {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  ColumnPtrs in;
  ARROW_RETURN_NOT_OK(in.Append(batches_, 0, /*allow_nulls=*/true));
  if (in.has_nulls) {  // cp::Sum skips nulls; the uint64 result of no valid row is 0 here (Arrow: null)
    ARROW_ASSIGN_OR_RAISE(b2_aggr_u32 a, Aggregates());
    return a.sum;
  }
  uint64_t sum = 0;
  b2_timings t{};
  b2_set* set = system_.set();  // contiguous batch ranges per GPU, one partial each (aggr_dpu.cc:82-84)
  B2_SET_RETURN_NOT_OK(set, b2_set_sum_u32_host(set, in.ptrs.data(), in.lens.data(),
                                                static_cast<int64_t>(in.ptrs.size()), &sum, &t));
  (void)ctx;
  timers_->Add(t);
  return sum;
}

}  // namespace aggr

// ---- Take (TakeDpu, host/take/take_dpu.cc:34-104) -----------------------------------------------
namespace take {

arrow::Status TakeGpu::Prepare() {
  timers_ = std::make_shared<timer::Timers>();
  return arrow::Status::OK();
}

arrow::Result<std::shared_ptr<arrow::Table>> TakeGpu::Run() {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  if (batches_.size() != indices_batches_.size())
    return arrow::Status::Invalid("values and indices must have the same number of batches");
  if (batches_.empty()) return arrow::Status::Invalid("no batches");
  ColumnPtrs v, i;
  ARROW_RETURN_NOT_OK(v.Append(batches_, 0, /*allow_nulls=*/true));
  ARROW_RETURN_NOT_OK(i.Append(indices_batches_, 0, /*allow_nulls=*/true));
  const int64_t nb = static_cast<int64_t>(v.ptrs.size());
  std::vector<uint32_t*> outs(nb);
  arrow::RecordBatchVector result;
  auto schema = batches_[0]->schema();
  int64_t total = 0;
  for (int64_t l : i.lens) total += l;
  ARROW_ASSIGN_OR_RAISE(auto slab, system_.pinned().Acquire(std::max<int64_t>(total, 1) * 4));
  int64_t off = 0;
  if (v.has_nulls || i.has_nulls) {
    // cp::Take semantics: slot j is null when index j is null or the value it selects is null; one
    // result batch per input batch over the slab, each with its own validity bitmap
    std::vector<uint8_t*> out_valid(nb);
    std::vector<std::shared_ptr<arrow::Buffer>> bitmaps(nb);
    for (int64_t b = 0; b < nb; ++b) {
      outs[b] = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(slab->data())) + off;
      ARROW_ASSIGN_OR_RAISE(bitmaps[b], arrow::AllocateBitmap(std::max<int64_t>(i.lens[b], 1)));
      out_valid[b] = bitmaps[b]->mutable_data();
      off += i.lens[b];
    }
    b2_timings t{};
    B2_ARROW_RETURN_NOT_OK(ctx, b2_take_u32_nullable_host(ctx, v.ptrs.data(), v.valid.data(), v.valid_off.data(),
                                                          v.lens.data(), i.ptrs.data(), i.valid.data(),
                                                          i.valid_off.data(), i.lens.data(), nb, outs.data(),
                                                          out_valid.data(), &t));
    timers_->Add(t);
    auto out_schema = arrow::schema({schema->field(0)->WithNullable(true)});
    off = 0;
    for (int64_t b = 0; b < nb; ++b) {
      auto arr = std::make_shared<arrow::UInt32Array>(i.lens[b], arrow::SliceBuffer(slab, off * 4, i.lens[b] * 4),
                                                      bitmaps[b], arrow::kUnknownNullCount);
      result.push_back(arrow::RecordBatch::Make(out_schema, i.lens[b], {std::move(arr)}));
      off += i.lens[b];
    }
    return arrow::Table::FromRecordBatches(out_schema, result);
  }
  for (int64_t b = 0; b < nb; ++b) {  // one result batch per input batch: zero-copy slices of the slab
    outs[b] = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(slab->data())) + off;
    auto arr = std::make_shared<arrow::UInt32Array>(i.lens[b], arrow::SliceBuffer(slab, off * 4, i.lens[b] * 4));
    result.push_back(arrow::RecordBatch::Make(schema, i.lens[b], {std::move(arr)}));
    off += i.lens[b];
  }
  b2_timings t{};
  b2_set* set = system_.set();  // batch-local gather: batch ranges per GPU
  B2_SET_RETURN_NOT_OK(set, b2_set_take_u32_host(set, v.ptrs.data(), v.lens.data(), i.ptrs.data(), i.lens.data(), nb,
                                                 outs.data(), &t));
  timers_->Add(t);
  return arrow::Table::FromRecordBatches(schema, result);
}

}  // namespace take

// ---- Join (JoinDpu, host/join/join_dpu.cc:144-400) ----------------------------------------------
namespace join {

arrow::Status JoinGpu::Prepare() {
  timers_ = std::make_shared<timer::Timers>();
  return arrow::Status::OK();
}

arrow::Result<std::shared_ptr<arrow::Table>> JoinGpu::Run() {
  b2_ctx* ctx = system_.ctx();
  b2_set* set = system_.set();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  const int fk = left_schema_->GetFieldIndex("fk"), pk = right_schema_->GetFieldIndex("pk");
  if (fk < 0 || pk < 0) return arrow::Status::Invalid("join keys are named fk / pk (join_native.cc:31-36)");
  // every non-key column of both sides comes along (join_dpu.cc:127-138,325-341)
  std::vector<int> lcols, rcols;
  for (int c = 0; c < left_schema_->num_fields(); ++c)
    if (c != fk) lcols.push_back(c);
  for (int c = 0; c < right_schema_->num_fields(); ++c)
    if (c != pk) rcols.push_back(c);
  ColumnPtrs l, r;  // [key batches..., payload 0 batches..., payload 1 batches..., ...]
  // key columns may carry nulls: a null key never matches (Arrow's hash join, join_native.cc:31-36)
  ARROW_RETURN_NOT_OK(l.Append(left_batches_, fk, /*allow_nulls=*/true));
  ARROW_RETURN_NOT_OK(r.Append(right_batches_, pk, /*allow_nulls=*/true));
  const bool null_keys = l.has_nulls || r.has_nulls;
  const std::vector<const uint8_t*> lkv = l.valid, rkv = r.valid;
  const std::vector<int64_t> lko = l.valid_off, rko = r.valid_off;
  for (int c : lcols) ARROW_RETURN_NOT_OK(l.Append(left_batches_, c));
  for (int c : rcols) ARROW_RETURN_NOT_OK(r.Append(right_batches_, c));
  const int64_t nlb = static_cast<int64_t>(left_batches_.size()), nrb = static_cast<int64_t>(right_batches_.size());
  // JoinNative drops pk and keeps fk + the payloads of both sides (join_native.cc:75)
  arrow::FieldVector fields{left_schema_->field(fk)};
  for (int c : lcols) fields.push_back(left_schema_->field(c));
  for (int c : rcols) fields.push_back(right_schema_->field(c));
  const int ncols = static_cast<int>(fields.size());
  uint64_t rows = 0;
  b2_timings t1{}, t2{};
  // phase timers: every member's local join is traced, the slowest member per phase is reported
  for (int i = 0; i < system_.size(); ++i) b2_join_trace(b2_set_ctx(set, i), 1);
  arrow::ArrayVector arrays(static_cast<size_t>(ncols));
  std::vector<uint32_t*> outs(static_cast<size_t>(ncols));
  const bool pair_path = lcols.size() == 1 && rcols.size() == 1;
  if (null_keys && !pair_path) return arrow::Status::NotImplemented("nullable join keys with one payload column per side");
  if (pair_path && null_keys) {
    // rows with a null key are dropped on the device before the join (member 0 of the set)
    B2_ARROW_RETURN_NOT_OK(ctx, b2_join_u32_nullable_host(ctx, l.ptrs.data(), lkv.data(), lko.data(), l.lens.data(), nlb,
                                                          r.ptrs.data(), rkv.data(), rko.data(), r.lens.data(), nrb,
                                                          &rows, &t1));
  } else if (pair_path) {
    // the benchmark shape: the payload travels with the key; sharded over every GPU of the set
    B2_SET_RETURN_NOT_OK(set, b2_set_join_u32_host(set, l.ptrs.data(), l.lens.data(), nlb, r.ptrs.data(),
                                                   r.lens.data(), nrb, &rows, &t1));
  } else {
    // row numbers + take kernel on member 0 (row numbers are local to one device's packed columns)
    B2_ARROW_RETURN_NOT_OK(ctx, b2_join_cols_u32_host(ctx, l.ptrs.data(), l.lens.data(), nlb,
                                                      static_cast<int>(lcols.size()), r.ptrs.data(), r.lens.data(),
                                                      nrb, static_cast<int>(rcols.size()), &rows, &t1));
  }
  for (int c = 0; c < ncols; ++c) {
    ARROW_ASSIGN_OR_RAISE(arrays[static_cast<size_t>(c)],
                          PinnedU32(system_, static_cast<int64_t>(rows), &outs[static_cast<size_t>(c)]));
  }
  if (pair_path && null_keys) {
    B2_ARROW_RETURN_NOT_OK(ctx, b2_join_fetch_host(ctx, outs[0], outs[1], outs[2], static_cast<int64_t>(rows), &t2));
  } else if (pair_path) {
    B2_SET_RETURN_NOT_OK(set, b2_set_join_fetch_host(set, outs[0], outs[1], outs[2], static_cast<int64_t>(rows), &t2));
  } else {
    B2_ARROW_RETURN_NOT_OK(ctx, b2_join_cols_fetch_host(ctx, outs.data(), ncols, static_cast<int64_t>(rows), &t2));
  }
  timers_->Add(t1);
  timers_->Add(t2);
  b2_join_phases slowest{};
  for (int i = 0; i < system_.size(); ++i) {
    b2_ctx* m = b2_set_ctx(set, i);
    b2_join_phases p{};
    B2_ARROW_RETURN_NOT_OK(m, b2_join_last_phases(m, &p));
    slowest.partition_build_ms = std::max(slowest.partition_build_ms, p.partition_build_ms);
    slowest.partition_probe_ms = std::max(slowest.partition_probe_ms, p.partition_probe_ms);
    slowest.probe_ms = std::max(slowest.probe_ms, p.probe_ms);
    slowest.take_ms = std::max(slowest.take_ms, p.take_ms);
    b2_join_trace(m, 0);
  }
  timers_->Add(slowest);
  return arrow::Table::Make(arrow::schema(fields), arrays, static_cast<int64_t>(rows));
}

// Fused pipeline [filter left payload < threshold ->] join -> COUNT / SUM(left payload) / SUM(right
// payload): one upload, three numbers back, nothing materialised (b2_join_aggr_u32_host).
arrow::Result<b2_join_aggr> JoinGpu::RunAggregate(bool filter_left_payload, uint32_t threshold) {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  const int fk = left_schema_->GetFieldIndex("fk"), pk = right_schema_->GetFieldIndex("pk");
  if (fk < 0 || pk < 0) return arrow::Status::Invalid("join keys are named fk / pk (join_native.cc:31-36)");
  if (left_schema_->num_fields() != 2 || right_schema_->num_fields() != 2)
    return arrow::Status::NotImplemented("one key and one payload column per side");
  ColumnPtrs l, r;
  ARROW_RETURN_NOT_OK(l.Append(left_batches_, fk));
  ARROW_RETURN_NOT_OK(l.Append(left_batches_, 1 - fk));
  ARROW_RETURN_NOT_OK(r.Append(right_batches_, pk));
  ARROW_RETURN_NOT_OK(r.Append(right_batches_, 1 - pk));
  b2_join_aggr out{};
  b2_timings t{};
  // sharded over every GPU of the set (one member: the single-context pipeline, filter pushed into the
  // probe side's first radix pass)
  b2_set* set = system_.set();
  (void)ctx;
  B2_SET_RETURN_NOT_OK(
      set, b2_set_join_aggr_u32_host(set, l.ptrs.data(), l.lens.data(), static_cast<int64_t>(left_batches_.size()),
                                     r.ptrs.data(), r.lens.data(), static_cast<int64_t>(right_batches_.size()),
                                     filter_left_payload ? 1 : 0, threshold, &out, &t));
  timers_->Add(t);
  return out;
}

}  // namespace join

// ---- Partition (PartitionDpu, host/partition/partition_dpu.cc:31-135) ---------------------------
namespace partition {

arrow::Status PartitionGpu::Prepare() {
  timers_ = std::make_shared<timer::Timers>();
  return arrow::Status::OK();
}

arrow::Result<arrow::RecordBatchVector> PartitionGpu::Run() {
  b2_ctx* ctx = system_.ctx();
  if (!timers_) timers_ = std::make_shared<timer::Timers>();
  const int ncols = schema_->num_fields();
  const int key = schema_->GetFieldIndex(partition_key_);
  if (key < 0) return arrow::Status::Invalid("no column named ", partition_key_);
  const int nparts = static_cast<int>(nr_partitions_);
  ColumnPtrs cols;  // column-major: all batches of column 0, then column 1, ...
  for (int c = 0; c < ncols; ++c) ARROW_RETURN_NOT_OK(cols.Append(batches_, c));
  std::vector<int64_t> lens;
  for (const auto& b : batches_) lens.push_back(b->num_rows());
  std::vector<int64_t> part_rows(nparts);
  b2_timings t1{}, t2{};
  B2_ARROW_RETURN_NOT_OK(
      ctx, b2_partition_u32_host(ctx, cols.ptrs.data(), lens.data(), static_cast<int64_t>(batches_.size()),
                                 ncols, key, nparts, part_rows.data(), &t1));
  std::vector<uint32_t*> outs(static_cast<size_t>(nparts) * ncols);
  arrow::RecordBatchVector result;
  for (int p = 0; p < nparts; ++p) {
    arrow::ArrayVector arrays;
    for (int c = 0; c < ncols; ++c) {
      ARROW_ASSIGN_OR_RAISE(auto arr, PinnedU32(system_, part_rows[p], &outs[static_cast<size_t>(p) * ncols + c]));
      arrays.push_back(std::move(arr));
    }
    result.push_back(arrow::RecordBatch::Make(schema_, part_rows[p], std::move(arrays)));
  }
  B2_ARROW_RETURN_NOT_OK(ctx, b2_partition_fetch_host(ctx, outs.data(), nparts, ncols, &t2));
  timers_->Add(t1);
  timers_->Add(t2);
  return result;
}

}  // namespace partition
}  // namespace upmemeval
