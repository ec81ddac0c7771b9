#include "native.h"

#include <arrow/acero/api.h>
#include <arrow/compute/api.h>
#include <arrow/compute/initialize.h>
#include <arrow/util/thread_pool.h>

namespace upmemeval {

namespace ac = ::arrow::acero;
namespace cp = ::arrow::compute;

arrow::Status InitNative(int threads) {
  ARROW_RETURN_NOT_OK(cp::Initialize());
  if (threads > 0) ARROW_RETURN_NOT_OK(arrow::SetCpuThreadPoolCapacity(threads));
  return arrow::Status::OK();
}

namespace filter {

arrow::Status FilterNative::Prepare() {
  ARROW_ASSIGN_OR_RAISE(input_, arrow::Table::FromRecordBatches(schema_, batches_));
  return arrow::Status::OK();
}

arrow::Result<std::shared_ptr<arrow::Table>> FilterNative::GetResult() {
  if (!input_) ARROW_RETURN_NOT_OK(Prepare());
  // source -> filter(v < 1<<30) -> sink   (filter_native.cc:52-66)
  cp::Expression pred = cp::less(cp::field_ref("v"), cp::literal(static_cast<uint32_t>(1u << 30)));
  ac::Declaration plan = ac::Declaration::Sequence(
      {{"table_source", ac::TableSourceNodeOptions{input_}}, {"filter", ac::FilterNodeOptions{pred}}});
  return ac::DeclarationToTable(std::move(plan), /*use_threads=*/true);
}

arrow::Result<uint64_t> FilterNative::Run() {
  ARROW_ASSIGN_OR_RAISE(auto t, GetResult());
  return static_cast<uint64_t>(t->num_rows());
}

}  // namespace filter

namespace aggr {

arrow::Status SumNative::Prepare() {
  ARROW_ASSIGN_OR_RAISE(input_, arrow::Table::FromRecordBatches(schema_, batches_));
  return arrow::Status::OK();
}

arrow::Result<uint64_t> SumNative::Run() {
  if (!input_) ARROW_RETURN_NOT_OK(Prepare());
  // source -> aggregate{"sum"(v)} -> sink   (aggr_native.cc:60-73)
  ac::Declaration plan = ac::Declaration::Sequence(
      {{"table_source", ac::TableSourceNodeOptions{input_}},
       {"aggregate", ac::AggregateNodeOptions{{{"sum", nullptr, "v", "sum(v)"}}}}});
  ARROW_ASSIGN_OR_RAISE(auto t, ac::DeclarationToTable(std::move(plan), /*use_threads=*/true));
  ARROW_ASSIGN_OR_RAISE(auto s, t->column(0)->GetScalar(0));
  return std::static_pointer_cast<arrow::UInt64Scalar>(s)->value;
}

}  // namespace aggr

namespace take {

arrow::Result<std::shared_ptr<arrow::Table>> TakeNative::Run() {
  // cp::Take per batch with NoBoundsCheck (take_native.cc:24-31)
  arrow::RecordBatchVector out;
  for (size_t b = 0; b < batches_.size(); ++b) {
    ARROW_ASSIGN_OR_RAISE(auto d, cp::Take(batches_[b]->column(0), indices_batches_[b]->column(0),
                                           cp::TakeOptions::NoBoundsCheck()));
    auto arr = d.make_array();
    out.push_back(arrow::RecordBatch::Make(schema_, arr->length(), {arr}));
  }
  return arrow::Table::FromRecordBatches(schema_, out);
}

}  // namespace take

namespace join {

arrow::Status JoinNative::Prepare() {
  ARROW_ASSIGN_OR_RAISE(left_, arrow::Table::FromRecordBatches(left_schema_, left_batches_));
  ARROW_ASSIGN_OR_RAISE(right_, arrow::Table::FromRecordBatches(right_schema_, right_batches_));
  return arrow::Status::OK();
}

arrow::Result<std::shared_ptr<arrow::Table>> JoinNative::Run() {
  if (!left_) ARROW_RETURN_NOT_OK(Prepare());
  // INNER hashjoin fk = pk, suffixes _l/_r, filter literal(true)   (join_native.cc:31-40)
  ac::HashJoinNodeOptions opts{ac::JoinType::INNER, {"fk"}, {"pk"}, cp::literal(true), "_l", "_r"};
  ac::Declaration l{"table_source", ac::TableSourceNodeOptions{left_}};
  ac::Declaration r{"table_source", ac::TableSourceNodeOptions{right_}};
  ac::Declaration plan{"hashjoin", {std::move(l), std::move(r)}, std::move(opts)};
  ARROW_ASSIGN_OR_RAISE(auto t, ac::DeclarationToTable(std::move(plan), /*use_threads=*/true));
  return t->RemoveColumn(t->schema()->GetFieldIndex("pk"));
}

}  // namespace join
}  // namespace upmemeval
