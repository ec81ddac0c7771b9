// host_test.cc — the reference's GoogleTests (host/filter/filter_test.cc, host/aggr/aggr_test.cc,
// host/take/take_test.cc, host/join/join_test.cc, host/partition/partition_test.cc) restated over
// the *Gpu operators: every case compares the device operator with the Native (Arrow Acero)
// operator on the same input. GoogleTest is not in the image, so a 30-line harness stands in.
// Exit code 0 iff every case passes. Needs a B200.
#include <arrow/api.h>
#include <arrow/compute/api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

#include "generator.h"
#include "native.h"
#include "operators.h"

using namespace upmemeval;
using namespace upmemeval::generator;

static int g_failed = 0;
#define EXPECT_TRUE(cond)                                                     \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::printf("    EXPECT_TRUE(%s) failed at %s:%d\n", #cond, __FILE__, __LINE__); \
      ++g_failed;                                                             \
    }                                                                         \
  } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

static std::shared_ptr<arrow::Schema> VSchema(const char* name = "v") {
  return arrow::schema({arrow::field(name, arrow::uint32(), /*nullable=*/false)});
}

// sort a table by the given columns (join_test.cc:27-38)
static std::shared_ptr<arrow::Table> Sorted(const std::shared_ptr<arrow::Table>& t,
                                            std::vector<std::string> keys) {
  std::vector<arrow::compute::SortKey> sk;
  for (auto& k : keys) sk.emplace_back(k);
  auto idx = arrow::compute::SortIndices(arrow::Datum(t), arrow::compute::SortOptions(sk)).ValueOrDie();
  return arrow::compute::Take(arrow::Datum(t), arrow::Datum(idx)).ValueOrDie().table()->CombineChunks().ValueOrDie();
}

// ---- FilterTest ------------------------------------------------------------------------------------
static void FilterSimpleTest(gpu::GpuSet& sys) {  // filter_test.cc:24-31
  auto rb = RecordBatchOf({"v"}, {ArrayOf({0, 2, 3, 8, 9})});
  filter::FilterGpu g{sys, {rb}};
  EXPECT_TRUE(g.Prepare().ok());
  filter::FilterNative n{rb->schema(), {rb}};
  EXPECT_TRUE(n.Prepare().ok());
  EXPECT_EQ(g.Run().ValueOrDie(), n.Run().ValueOrDie());
}
static void FilterResultTest(gpu::GpuSet& sys) {  // filter_test.cc:33-61
  std::vector<uint32_t> v(4096);
  for (int i = 0; i < 4096; ++i) v[i] = (i == 5 || i == 8 || i == 9 || i == 100 || i == 270) ? i : i + (1u << 30);
  auto rb = RecordBatchOf({"v"}, {ArrayOf(v)});
  filter::FilterGpu g{sys, {rb}};
  EXPECT_TRUE(g.Prepare().ok());
  auto gr = g.GetResult().ValueOrDie();
  filter::FilterNative n{rb->schema(), {rb}};
  auto nr = n.GetResult().ValueOrDie()->column(0);
  EXPECT_TRUE(gr->Equals(nr));
  EXPECT_EQ(gr->length(), 5);
}
static void FilterLongerTest(gpu::GpuSet& sys) {  // filter_test.cc:63-78, plus a multi-batch case
  for (int nb : {1, 128}) {
    RandomArrayGenerator rng(42);
    auto schema = VSchema();
    auto batches = MakeRandomRecordBatches(rng, schema, nb, 1 << 16);
    filter::FilterGpu g{sys, batches};
    EXPECT_TRUE(g.Prepare().ok());
    auto gr = g.GetResult().ValueOrDie();
    if (nb == 1) {
      filter::FilterNative n{schema, batches};
      EXPECT_TRUE(gr->Equals(n.GetResult().ValueOrDie()->column(0)));
      EXPECT_EQ(gr->length(), 16358);  // SURVEY.md §8(c) fingerprint of generator(42)
    } else {  // Acero does not keep batch order with threads: compare chunk by chunk instead
      EXPECT_EQ(gr->num_chunks(), nb);
      for (int b = 0; b < nb; ++b) {
        filter::FilterNative n{schema, {batches[b]}};
        EXPECT_TRUE(arrow::ChunkedArray(gr->chunk(b)).Equals(*n.GetResult().ValueOrDie()->column(0)));
      }
    }
  }
}

static void FilterPinnedGatherTest(gpu::GpuSet& sys) {  // same inputs, page-locked: gather-kernel upload
  RandomArrayGenerator rng(7);
  auto schema = VSchema();
  auto batches = MakeRandomRecordBatches(rng, schema, 300, 1 << 16);
  filter::FilterGpu plain{sys, batches};
  auto expect = plain.GetResult().ValueOrDie();
  {
    gpu::PinnedBatches pinned(batches);
    EXPECT_EQ(pinned.regions(), 300u);
    sys.PromiseInputsPinned(true);
    filter::FilterGpu g{sys, batches};
    EXPECT_TRUE(g.GetResult().ValueOrDie()->Equals(expect));
    aggr::SumGpu s{sys, batches};
    aggr::SumNative n{schema, batches};
    EXPECT_EQ(s.Run().ValueOrDie(), n.Run().ValueOrDie());
    sys.PromiseInputsPinned(false);
  }
}

// ---- SumTest ----------------------------------------------------------------------------------------
static void SumSimpleTest(gpu::GpuSet& sys) {  // aggr_test.cc:24-36
  auto rb = RecordBatchOf({"v"}, {ArrayOf({0, 2, 3, 8, 9})});
  aggr::SumGpu g{sys, {rb}};
  EXPECT_TRUE(g.Prepare().ok());
  EXPECT_EQ(g.Run().ValueOrDie(), 22u);
  aggr::SumNative n{rb->schema(), {rb}};
  EXPECT_EQ(n.Run().ValueOrDie(), 22u);
}
static void SumLargeTest(gpu::GpuSet& sys) {  // aggr_test.cc:38-49
  RandomArrayGenerator rng(42);
  auto schema = VSchema();
  auto batches = MakeRandomRecordBatches(rng, schema, 128, 1 << 16);
  aggr::SumGpu g{sys, batches};
  EXPECT_TRUE(g.Prepare().ok());
  aggr::SumNative n{schema, batches};
  EXPECT_EQ(g.Run().ValueOrDie(), n.Run().ValueOrDie());
}

// ---- TakeTest ---------------------------------------------------------------------------------------
static void TakeSimpleTest(gpu::GpuSet& sys) {  // take_test.cc:24-46
  auto rb = RecordBatchOf({"v"}, {ArrayOf({0, 2, 3, 8, 9})});
  auto idx = RecordBatchOf({"i"}, {ArrayOf({0, 1, 4})});
  take::TakeGpu g{sys, {rb}, {idx}};
  EXPECT_TRUE(g.Prepare().ok());
  auto t = g.Run().ValueOrDie();
  auto a = std::static_pointer_cast<arrow::UInt32Array>(t->column(0)->chunk(0));
  EXPECT_EQ(a->length(), 3);
  EXPECT_EQ(a->Value(0), 0u);
  EXPECT_EQ(a->Value(1), 2u);
  EXPECT_EQ(a->Value(2), 9u);
}
static void TakeLargeTest(gpu::GpuSet& sys) {  // take_test.cc:48-72
  const int nb = 128, bs = 64 << 10, ibs = 8 << 10;
  RandomArrayGenerator rng(42);
  auto schema = VSchema();
  auto batches = MakeRandomRecordBatches(rng, schema, nb, bs);
  auto md = arrow::key_value_metadata({{"min", "0"}, {"max", std::to_string(bs - 1)}});
  auto ischema = arrow::schema({arrow::field("i", arrow::uint32(), false, md)});
  auto indices = MakeRandomRecordBatches(rng, ischema, nb, ibs);
  take::TakeGpu g{sys, batches, indices};
  EXPECT_TRUE(g.Prepare().ok());
  take::TakeNative n{schema, batches, indices};
  EXPECT_TRUE(n.Run().ValueOrDie()->Equals(*g.Run().ValueOrDie()));
}

// ---- nullable columns: the Gpu operators against the Native (Arrow) ones ------
// Batches of `rows` uint32 with every `null_every`-th row null; slice_off > 0 slices them so the
// validity bitmap carries a bit offset.
static arrow::RecordBatchVector NullableBatches(const char* name, int nb, int rows, int null_every,
                                                uint32_t max_value, int slice_off, uint32_t seed) {
  arrow::RecordBatchVector out;
  auto schema = arrow::schema({arrow::field(name, arrow::uint32(), /*nullable=*/true)});
  uint32_t x = seed;
  for (int b = 0; b < nb; ++b) {
    arrow::UInt32Builder bld;
    for (int r = 0; r < rows + slice_off; ++r) {
      x = x * 1664525u + 1013904223u;
      if (null_every > 0 && (x >> 8) % null_every == 0) (void)bld.AppendNull();
      else (void)bld.Append(max_value == 0xffffffffu ? x : (x >> 4) % (max_value + 1));
    }
    auto arr = bld.Finish().ValueOrDie()->Slice(slice_off, rows);
    out.push_back(arrow::RecordBatch::Make(schema, rows, {arr}));
  }
  return out;
}
static void FilterNullableTest(gpu::GpuSet& sys) {
  for (int slice_off : {0, 5}) {
    auto batches = NullableBatches("v", 6, 8192, 3, 0xffffffffu, slice_off, 7);
    filter::FilterGpu g{sys, batches};
    EXPECT_TRUE(g.Prepare().ok());
    auto gr = g.GetResult().ValueOrDie();
    filter::FilterNative n{batches[0]->schema(), batches};
    auto nr = n.GetResult().ValueOrDie()->column(0);
    EXPECT_TRUE(gr->Equals(nr));
    EXPECT_EQ(gr->null_count(), 0);
    EXPECT_TRUE(gr->length() > 0 && gr->length() < 6 * 8192);
  }
}
// Other fixed-width column types (the reference fixes T = uint32_t): the typed
// FilterGpu against Arrow's filter(less(column, threshold)), batch by batch, with nulls and a sliced
// batch (bit offset in the validity bitmap), NaN and infinities in the float columns.
template <typename ArrowType>
static void FilterTypedCase(gpu::GpuSet& sys, std::shared_ptr<arrow::Scalar> threshold, uint32_t seed) {
  using CType = typename ArrowType::c_type;
  auto type = arrow::TypeTraits<ArrowType>::type_singleton();
  auto schema = arrow::schema({arrow::field("v", type, /*nullable=*/true)});
  arrow::RecordBatchVector batches;
  uint64_t x = seed;
  const int lens[] = {65536, 1, 0, 4097, 30000};
  for (int rows : lens) {
    typename arrow::TypeTraits<ArrowType>::BuilderType bld;
    for (int r = 0; r < rows + 3; ++r) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t h = x ^ (x >> 29);
      if ((h >> 7) % 5 == 0) { (void)bld.AppendNull(); continue; }
      CType v;
      if (std::is_floating_point<CType>::value) {
        const int k = (int)((h >> 11) % 64);
        v = k == 0 ? std::numeric_limits<CType>::quiet_NaN()
            : k == 1 ? std::numeric_limits<CType>::infinity()
            : k == 2 ? -std::numeric_limits<CType>::infinity()
                     : (CType)((double)(int64_t)(h >> 20) / 1e9 - 8000.0);
      } else {
        std::memcpy(&v, &h, sizeof(CType));
      }
      (void)bld.Append(v);
    }
    auto arr = bld.Finish().ValueOrDie()->Slice(3, rows);
    batches.push_back(arrow::RecordBatch::Make(schema, rows, {arr}));
  }
  filter::FilterGpu g{sys, batches, threshold};
  EXPECT_TRUE(g.Prepare().ok());
  auto gr = g.GetResult().ValueOrDie();
  EXPECT_EQ(gr->num_chunks(), (int)batches.size());
  EXPECT_TRUE(gr->type()->Equals(*type));
  int64_t selected = 0;
  for (size_t b = 0; b < batches.size(); ++b) {
    auto col = batches[b]->column(0);
    auto thr = arrow::compute::Cast(arrow::Datum(threshold), type).ValueOrDie();
    auto mask = arrow::compute::CallFunction("less", {arrow::Datum(col), thr}).ValueOrDie();
    auto want = arrow::compute::Filter(arrow::Datum(col), mask).ValueOrDie().make_array();
    EXPECT_EQ(want->null_count(), 0);
    // bit patterns, not values: NaN never passes, -0.0 stays -0.0
    auto got = gr->chunk((int)b);
    EXPECT_EQ(got->length(), want->length());
    if (got->length() == want->length() && got->length() > 0)
      EXPECT_TRUE(std::memcmp(got->data()->template GetValues<CType>(1), want->data()->template GetValues<CType>(1),
                              (size_t)got->length() * sizeof(CType)) == 0);
    selected += want->length();
  }
  EXPECT_EQ((int64_t)g.Run().ValueOrDie(), selected);
  EXPECT_TRUE(selected > 0);
}
static void FilterTypedTest(gpu::GpuSet& sys) {
  FilterTypedCase<arrow::Int32Type>(sys, arrow::MakeScalar((int32_t)-12345), 1);
  FilterTypedCase<arrow::FloatType>(sys, arrow::MakeScalar(0.5f), 2);
  FilterTypedCase<arrow::UInt64Type>(sys, arrow::MakeScalar((uint64_t)1 << 62), 3);
  FilterTypedCase<arrow::Int64Type>(sys, arrow::MakeScalar((int64_t)-(1ll << 40)), 4);
  FilterTypedCase<arrow::DoubleType>(sys, arrow::MakeScalar(250.0), 5);
  FilterTypedCase<arrow::Int64Type>(sys, arrow::MakeScalar((uint32_t)(1u << 30)), 6);  // threshold cast to the column's type
}
static void SumNullableTest(gpu::GpuSet& sys) {
  auto batches = NullableBatches("v", 5, 50000, 4, 0xffffffffu, 3, 11);
  aggr::SumGpu g{sys, batches};
  EXPECT_TRUE(g.Prepare().ok());
  aggr::SumNative n{batches[0]->schema(), batches};
  EXPECT_EQ(g.Run().ValueOrDie(), n.Run().ValueOrDie());
  auto table = arrow::Table::FromRecordBatches(batches).ValueOrDie();
  auto mm = arrow::compute::MinMax(table->column(0)).ValueOrDie().scalar_as<arrow::StructScalar>();
  const b2_aggr_u32 a = g.Aggregates().ValueOrDie();
  EXPECT_EQ(a.count, (uint64_t)(table->num_rows() - table->column(0)->null_count()));
  EXPECT_EQ(a.min, std::static_pointer_cast<arrow::UInt32Scalar>(mm.value[0])->value);
  EXPECT_EQ(a.max, std::static_pointer_cast<arrow::UInt32Scalar>(mm.value[1])->value);
}
static void TakeNullableTest(gpu::GpuSet& sys) {
  auto values = NullableBatches("v", 4, 4096, 3, 0xffffffffu, 2, 13);
  auto indices = NullableBatches("i", 4, 1000, 5, 4095, 1, 17);
  take::TakeGpu g{sys, values, indices};
  EXPECT_TRUE(g.Prepare().ok());
  auto gt = g.Run().ValueOrDie();
  take::TakeNative n{values[0]->schema(), values, indices};
  auto nt = n.Run().ValueOrDie();
  EXPECT_TRUE(gt->column(0)->Equals(nt->column(0)));
  EXPECT_TRUE(gt->column(0)->null_count() > 0);
  // the join has no null semantics here: rejected loudly
  auto l = NullableBatches("fk", 1, 10, 2, 100, 0, 1);
  join::JoinGpu j{sys, l[0]->schema(), l[0]->schema(), l, l};
  EXPECT_TRUE(!j.Run().ok());
}

// ---- JoinTest ---------------------------------------------------------------------------------------
static void JoinSimpleTest(gpu::GpuSet& sys) {  // join_test.cc:40-80
  arrow::RecordBatchVector left = {
      RecordBatchOf({"fk", "v"}, {ArrayOf({0, 2, 3, 8, 9}), ArrayOf({100, 102, 103, 108, 109})}),
      RecordBatchOf({"fk", "v"}, {ArrayOf({10, 12, 13, 18, 19}), ArrayOf({110, 112, 113, 118, 119})})};
  arrow::RecordBatchVector right = {
      RecordBatchOf({"pk", "v"}, {ArrayOf({3, 8, 9, 0, 2}), ArrayOf({53, 58, 59, 50, 52})}),
      RecordBatchOf({"pk", "v"}, {ArrayOf({12, 13, 18, 19, 10}), ArrayOf({62, 63, 68, 69, 60})})};
  join::JoinGpu g{sys, left[0]->schema(), right[0]->schema(), left, right};
  EXPECT_TRUE(g.Prepare().ok());
  auto gt = g.Run().ValueOrDie();
  EXPECT_EQ(gt->num_rows(), 10);
  join::JoinNative n{left[0]->schema(), right[0]->schema(), left, right};
  auto nt = n.Run().ValueOrDie();
  // Acero names the clashing payload columns v_l / v_r; compare positionally after sorting
  auto gs = Sorted(gt->RenameColumns({"fk", "a", "b"}).ValueOrDie(), {"a", "fk"});
  auto ns = Sorted(nt->RenameColumns({"fk", "a", "b"}).ValueOrDie(), {"a", "fk"});
  for (int c = 0; c < 3; ++c) EXPECT_TRUE(gs->column(c)->Equals(ns->column(c)));
}
static void JoinLargeTest(gpu::GpuSet& sys) {  // join_test.cc:82-121 (128 x 65536 per side)
  const int nb = 128, bs = 64 << 10;
  RandomArrayGenerator rng(42);
  auto rschema0 = VSchema("x"), lschema0 = VSchema("y");
  auto right = AddColumn("pk", MakeRandomRecordBatches(rng, rschema0, nb, bs), MakeIndexColumn(nb, bs).ValueOrDie());
  auto lefty = MakeRandomRecordBatches(rng, lschema0, nb, bs);
  auto left = AddColumn("fk", lefty, MakeForeignKeyColumn(rng, bs, nb, bs).ValueOrDie());
  join::JoinGpu g{sys, left[0]->schema(), right[0]->schema(), left, right};
  EXPECT_TRUE(g.Prepare().ok());
  auto gt = g.Run().ValueOrDie();
  EXPECT_EQ(gt->num_rows(), static_cast<int64_t>(nb) * bs);  // :115-116
  // JoinDpu's phase timers (join_dpu.cc:146-148) come back through the ABI (b2_join_last_phases)
  const auto& timers = g.Timers()->get();
  for (const char* name : {"partitionKernel", "probe"}) {
    auto it = timers.find(name);
    EXPECT_TRUE(it != timers.end() && it->second->Result().count() > 0);
  }
  EXPECT_TRUE(timers.count("take") == 1);
  join::JoinNative n{left[0]->schema(), right[0]->schema(), left, right};
  auto nt = n.Run().ValueOrDie();
  auto gs = Sorted(gt, {"fk", "y"}), ns = Sorted(nt, {"fk", "y"});
  for (const char* c : {"fk", "y", "x"}) EXPECT_TRUE(gs->GetColumnByName(c)->Equals(ns->GetColumnByName(c)));
}

static void JoinManyPayloadsTest(gpu::GpuSet& sys) {  // JoinDpu carries every non-key column (join_dpu.cc:127-138,325-341)
  const int nb = 8, bs = 64 << 10;
  RandomArrayGenerator rng(42);
  auto right = AddColumn("pk", MakeRandomRecordBatches(rng, VSchema("x"), nb, bs), MakeIndexColumn(nb, bs).ValueOrDie());
  auto x2 = MakeRandomRecordBatches(rng, VSchema("x2"), nb, bs);
  for (int b = 0; b < nb; ++b) right[b] = right[b]->AddColumn(2, "x2", x2[b]->column(0)).ValueOrDie();
  auto left = AddColumn("fk", MakeRandomRecordBatches(rng, VSchema("y"), nb, bs),
                        MakeForeignKeyColumn(rng, bs, nb, bs).ValueOrDie());
  auto y2 = MakeRandomRecordBatches(rng, VSchema("y2"), nb, bs);
  auto y3 = MakeRandomRecordBatches(rng, VSchema("y3"), nb, bs);
  for (int b = 0; b < nb; ++b) {
    left[b] = left[b]->AddColumn(2, "y2", y2[b]->column(0)).ValueOrDie();
    left[b] = left[b]->AddColumn(3, "y3", y3[b]->column(0)).ValueOrDie();
  }
  join::JoinGpu g{sys, left[0]->schema(), right[0]->schema(), left, right};
  EXPECT_TRUE(g.Prepare().ok());
  auto gt = g.Run().ValueOrDie();
  EXPECT_EQ(gt->num_rows(), static_cast<int64_t>(nb) * bs);
  EXPECT_EQ(gt->num_columns(), 6);
  join::JoinNative n{left[0]->schema(), right[0]->schema(), left, right};
  auto nt = n.Run().ValueOrDie();
  auto gs = Sorted(gt, {"fk", "y", "y2"}), ns = Sorted(nt, {"fk", "y", "y2"});
  for (const char* c : {"fk", "y", "y2", "y3", "x", "x2"})
    EXPECT_TRUE(gs->GetColumnByName(c)->Equals(ns->GetColumnByName(c)));
}

static void JoinAggregateTest(gpu::GpuSet& sys) {  // fused pipeline vs aggregates of the Native join's columns
  RandomArrayGenerator rng(42);
  const int nb = 16, bs = 1 << 16;
  auto xs = MakeRandomRecordBatches(rng, VSchema("x"), nb, bs);
  auto ys = MakeRandomRecordBatches(rng, VSchema("y"), nb, bs);
  auto fks = MakeForeignKeyColumn(rng, bs, nb, bs).ValueOrDie();
  auto pks = MakeIndexColumn(nb, bs).ValueOrDie();
  arrow::RecordBatchVector left, right;
  for (int b = 0; b < nb; ++b) {
    left.push_back(RecordBatchOf({"fk", "y"}, {fks[b], ys[b]->column(0)}));
    right.push_back(RecordBatchOf({"pk", "x"}, {pks[b], xs[b]->column(0)}));
  }
  join::JoinGpu g{sys, left[0]->schema(), right[0]->schema(), left, right};
  EXPECT_TRUE(g.Prepare().ok());
  join::JoinNative n{left[0]->schema(), right[0]->schema(), left, right};
  auto t = n.Run().ValueOrDie();
  auto sum_of = [&](const char* name) {
    auto col = t->GetColumnByName(name);
    auto c = arrow::compute::Cast(arrow::Datum(col), arrow::uint64()).ValueOrDie();
    return std::static_pointer_cast<arrow::UInt64Scalar>(arrow::compute::Sum(c).ValueOrDie().scalar())->value;
  };
  const b2_join_aggr a = g.RunAggregate().ValueOrDie();
  EXPECT_EQ(a.rows, (uint64_t)t->num_rows());
  EXPECT_EQ(a.sum_y, sum_of("y"));
  EXPECT_EQ(a.sum_x, sum_of("x"));
  // with the pushed-down filter: the Native join over the pre-filtered probe side
  const uint32_t thr = 1u << 30;
  auto mask = arrow::compute::CallFunction("less", {arrow::Datum(t->GetColumnByName("y")),
                                                    arrow::Datum(std::make_shared<arrow::UInt32Scalar>(thr))}).ValueOrDie();
  auto ft = arrow::compute::Filter(arrow::Datum(t), mask).ValueOrDie().table();
  const b2_join_aggr f = g.RunAggregate(true, thr).ValueOrDie();
  EXPECT_EQ(f.rows, (uint64_t)ft->num_rows());
  EXPECT_TRUE(f.rows > 0 && f.rows < a.rows);
}

// ---- PartitionTest (GTEST_SKIP in the reference) ------------------------------------------------------
static void PartitionSimpleTest(gpu::GpuSet& sys) {  // partition_test.cc:21-57
  arrow::RecordBatchVector batches = {RecordBatchOf({"pk", "x"}, {ArrayOf({0, 2}), ArrayOf({100, 101})}),
                                      RecordBatchOf({"pk", "x"}, {ArrayOf({3, 8}), ArrayOf({102, 103})})};
  partition::PartitionGpu p{sys, batches[0]->schema(), batches, 2, "pk"};
  EXPECT_TRUE(p.Prepare().ok());
  auto parts = p.Run().ValueOrDie();
  EXPECT_EQ(parts.size(), 2u);
  int64_t sizes[2] = {parts[0]->num_rows(), parts[1]->num_rows()};
  EXPECT_TRUE((sizes[0] == 3 && sizes[1] == 1) || (sizes[0] == 1 && sizes[1] == 3));
  uint64_t spk = 0, sx = 0;
  for (auto& b : parts)
    for (int64_t i = 0; i < b->num_rows(); ++i) {
      spk += std::static_pointer_cast<arrow::UInt32Array>(b->column(0))->Value(i);
      sx += std::static_pointer_cast<arrow::UInt32Array>(b->column(1))->Value(i);
    }
  EXPECT_EQ(spk, 13u);
  EXPECT_EQ(sx, 406u);
}
static void PartitionLargeTest(gpu::GpuSet& sys) {  // partition_test.cc:59-92
  const int nb = 128, bs = 64 << 10, nparts = 32;
  RandomArrayGenerator rng(42);
  auto batches = AddColumn("pk", MakeRandomRecordBatches(rng, VSchema("x"), nb, bs), MakeIndexColumn(nb, bs).ValueOrDie());
  partition::PartitionGpu p{sys, batches[0]->schema(), batches, nparts, "pk"};
  EXPECT_TRUE(p.Prepare().ok());
  auto parts = p.Run().ValueOrDie();
  const double mean = static_cast<double>(nb) * bs / nparts;
  int64_t total = 0;
  for (auto& b : parts) {
    total += b->num_rows();
    EXPECT_TRUE(std::abs(b->num_rows() - mean) < 0.1 * mean);
    auto pk = std::static_pointer_cast<arrow::UInt32Array>(b->column(0));
    for (int64_t i = 0; i < b->num_rows(); i += 997)
      EXPECT_EQ(static_cast<int>(b2_wang_hash_u32(pk->Value(i)) >> 27), static_cast<int>(&b - &parts[0]));
  }
  EXPECT_EQ(total, static_cast<int64_t>(nb) * bs);
}

// ---- CPU-only cases (--cpu): the generator restatement and the Native plans ---------------------------
static int RunCpuCases() {
  RandomArrayGenerator rng(42);
  auto schema = VSchema();
  auto batches = MakeRandomRecordBatches(rng, schema, 1, 1 << 16);
  auto v = std::static_pointer_cast<arrow::UInt32Array>(batches[0]->column(0));
  // fingerprints of RandomArrayGenerator(42) produced by the real libstdc++/PCG headers
  // (tests/golden/generator_golden.json)
  EXPECT_EQ(v->Value(0), 268u);
  EXPECT_EQ(v->Value(1), 2955549055u);
  EXPECT_EQ(v->Value(2), 1465994917u);
  uint64_t sum = 0, cnt = 0;
  for (int64_t i = 0; i < v->length(); ++i) {
    sum += v->Value(i);
    cnt += v->Value(i) < (1u << 30);
  }
  EXPECT_EQ(sum, 141101534903199ull);
  EXPECT_EQ(cnt, 16358u);
  filter::FilterNative f{schema, batches};
  EXPECT_EQ(f.Run().ValueOrDie(), 16358u);
  aggr::SumNative s{schema, batches};
  EXPECT_EQ(s.Run().ValueOrDie(), 141101534903199ull);
  auto rb = RecordBatchOf({"v"}, {ArrayOf({0, 2, 3, 8, 9})});
  auto idx = RecordBatchOf({"i"}, {ArrayOf({0, 1, 4})});
  take::TakeNative t{rb->schema(), {rb}, {idx}};
  auto ta = std::static_pointer_cast<arrow::UInt32Array>(t.Run().ValueOrDie()->column(0)->chunk(0));
  EXPECT_TRUE(ta->length() == 3 && ta->Value(0) == 0 && ta->Value(1) == 2 && ta->Value(2) == 9);
  arrow::RecordBatchVector left = {RecordBatchOf({"fk", "y"}, {ArrayOf({0, 2, 3, 8, 9}), ArrayOf({100, 102, 103, 108, 109})})};
  arrow::RecordBatchVector right = {RecordBatchOf({"pk", "x"}, {ArrayOf({3, 8, 9, 0, 2, 7}), ArrayOf({53, 58, 59, 50, 52, 57})})};
  join::JoinNative j{left[0]->schema(), right[0]->schema(), left, right};
  auto jt = Sorted(j.Run().ValueOrDie(), {"fk"});
  EXPECT_EQ(jt->num_rows(), 5);
  EXPECT_EQ(jt->schema()->field_names(), (std::vector<std::string>{"fk", "y", "x"}));
  EXPECT_EQ(std::static_pointer_cast<arrow::UInt32Array>(jt->column(2)->chunk(0))->Value(4), 59u);
  // the pk counter and the fk ranges of the generator (generator.cc:46-71)
  auto pk = MakeIndexColumn(2, 4).ValueOrDie();
  EXPECT_EQ(std::static_pointer_cast<arrow::UInt32Array>(pk[1])->Value(3), 7u);
  auto fk = MakeForeignKeyColumn(rng, 1 << 21, 3, 1000).ValueOrDie();
  for (int b = 0; b < 3; ++b) {
    auto a = std::static_pointer_cast<arrow::UInt32Array>(fk[b]);
    for (int64_t i = 0; i < a->length(); ++i)
      EXPECT_TRUE(a->Value(i) >= (uint32_t)b << 21 && a->Value(i) < (uint32_t)(b + 1) << 21);
  }
  std::printf("[%s] cpu cases (generator fingerprints, Native known answers)\n", g_failed ? "FAILED" : "  OK  ");
  return g_failed ? 1 : 0;
}

int main(int argc, char** argv) {
  if (!InitNative(0).ok()) return 2;
  if (argc > 1 && std::string(argv[1]) == "--cpu") return RunCpuCases();
  const char* ng = std::getenv("NR_GPUS");  // the same cases over a set of several GPUs
  auto sys = gpu::GpuSet::allocate(ng ? std::atoi(ng) : 1, 0);
  if (!sys.ok()) {
    std::printf("no GPU: %s\n", sys.status().ToString().c_str());
    return 3;
  }
  struct Case { const char* name; std::function<void(gpu::GpuSet&)> fn; };
  std::vector<Case> cases = {
      {"FilterTest.SimpleTest", FilterSimpleTest}, {"FilterTest.ResultTest", FilterResultTest},
      {"FilterTest.LongerTest", FilterLongerTest}, {"FilterTest.PinnedGather", FilterPinnedGatherTest},
      {"SumTest.SimpleTest", SumSimpleTest},
      {"SumTest.LargeTest", SumLargeTest},         {"TakeTest.SimpleTest", TakeSimpleTest},
      {"TakeTest.LargeTest", TakeLargeTest},       {"FilterTest.Nullable", FilterNullableTest},
      {"FilterTest.TypedColumns", FilterTypedTest},
      {"SumTest.Nullable", SumNullableTest},       {"TakeTest.Nullable", TakeNullableTest},
      {"JoinTest.SimpleTest", JoinSimpleTest},
      {"JoinTest.LargeTest", JoinLargeTest},       {"JoinTest.ManyPayloads", JoinManyPayloadsTest},
      {"JoinTest.FusedAggregate", JoinAggregateTest},       {"PartitionTest.SimpleTest", PartitionSimpleTest},
      {"PartitionTest.LargeTest", PartitionLargeTest}};
  int bad = 0;
  for (auto& c : cases) {
    const int before = g_failed;
    c.fn(**sys);
    std::printf("[%s] %s\n", g_failed == before ? "  OK  " : "FAILED", c.name);
    bad += g_failed != before;
  }
  std::printf("%d of %zu cases failed\n", bad, cases.size());
  return bad ? 1 : 0;
}
