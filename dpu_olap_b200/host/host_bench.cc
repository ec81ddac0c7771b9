// host_bench.cc — the reference's benchmark harness (host/main_benchmark.cc:6-17 and the BM_*
// fixtures in host/*/..._benchmark.cc) with BM_*Gpu cases beside the BM_*Native (Arrow Acero)
// ones on identical generator(42) inputs. Google Benchmark is not in the image, so this driver
// times Prepare()+Run() itself (as BM_Filter does, filter_benchmark.cc:30-49) and prints the same
// JSON shape scripts/parse_results.py:25-35 reads: benchmarks[].name =
// "<Fixture>/<BM_Name>/<Arg>:<v>/...", real_time, counters, error_occurred.
//
//   SF=64 MAX_THREADS=16 ./host_bench [--benchmark_filter=BM_Filter] [--iterations=3]
#include <arrow/api.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "generator.h"
#include "native.h"
#include "operators.h"

using namespace upmemeval;
using namespace upmemeval::generator;
using Clock = std::chrono::steady_clock;

static int EnvInt(const char* name, int dflt) {  // host/system/system.h:7-20
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

struct Result {
  std::string name;
  double real_ms = 0;
  int iterations = 0;
  bool error = false;
  std::string error_message;
  std::map<std::string, double> counters;
};

template <typename Fn>
static Result TimeIt(const std::string& name, int iters, double items, double bytes, Fn&& fn) {
  Result r;
  r.name = name;
  double total = 0;
  std::vector<double> times;
  for (int i = 0; i < iters + 1; ++i) {  // first iteration is a warm-up
    const auto t0 = Clock::now();
    arrow::Status st = fn(r);
    const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
    if (!st.ok()) {
      r.error = true;
      r.error_message = st.ToString();
      return r;
    }
    if (i > 0) {
      total += ms;
      times.push_back(ms);
    }
  }
  r.iterations = iters;
  r.real_ms = total / iters;
  std::sort(times.begin(), times.end());
  r.counters["real_time_min_ms"] = times.front();
  r.counters["real_time_median_ms"] = times[times.size() / 2];
  r.counters["items_per_second"] = items / (r.real_ms * 1e-3);
  r.counters["bytes_per_second"] = bytes / (r.real_ms * 1e-3);
  return r;
}

static void AddTimers(Result& r, const std::shared_ptr<timer::Timers>& t) {
  if (!t) return;
  for (auto& kv : t->get()) r.counters[kv.first] = kv.second->Result().count() / 1e6;  // ms
}

int main(int argc, char** argv) {
  std::string filter, out_path;
  int iters = 3;
  for (int i = 1; i < argc; ++i) {  // the Google Benchmark flags the reference's runners pass
    if (!std::strncmp(argv[i], "--benchmark_filter=", 19)) filter = argv[i] + 19;
    if (!std::strncmp(argv[i], "--benchmark_out=", 16)) out_path = argv[i] + 16;
    if (!std::strncmp(argv[i], "--benchmark_repetitions=", 24)) iters = std::atoi(argv[i] + 24);
    if (!std::strncmp(argv[i], "--iterations=", 13)) iters = std::atoi(argv[i] + 13);
  }
  if (iters < 1) iters = 1;
  const int sf = EnvInt("SF", 1);
  const int threads = EnvInt("MAX_THREADS", static_cast<int>(std::thread::hardware_concurrency()));
  if (!InitNative(threads).ok()) return 2;
  // NR_GPUS: the whole set, as NR_DPUS in the reference (host/system/system.h:14-20, main_benchmark.cc:7-8)
  auto sys_r = gpu::GpuSet::allocate(EnvInt("NR_GPUS", 1), EnvInt("GPU", 0));
  gpu::GpuSet* sys = sys_r.ok() ? sys_r->get() : nullptr;
  if (sys) sys->PromiseInputsPinned(true);  // every fixture below pins its batches (gpu::PinnedBatches)
  auto want = [&](const char* n) { return filter.empty() || std::strstr(n, filter.c_str()) != nullptr; };
  std::vector<Result> results;
  auto vschema = [](const char* n) { return arrow::schema({arrow::field(n, arrow::uint32(), false)}); };

  // ---- filter: SF<<7 batches of 64 Ki rows (filter_benchmark.cc:142,153) ----
  if (want("BM_FilterNative") || want("BM_FilterGpu")) {
    RandomArrayGenerator rng(42);
    const int nb = sf << 7, bs = 64 << 10;
    auto schema = vschema("v");
    auto batches = MakeRandomRecordBatches(rng, schema, nb, bs);
    gpu::PinnedBatches pinned(batches);  // fixture set-up (untimed): page-lock the Arrow buffers in place
    const double rows = static_cast<double>(nb) * bs;
    const std::string args = "/Batches:" + std::to_string(nb) + "/BatchSize:" + std::to_string(bs);
    if (want("BM_FilterNative"))
      results.push_back(TimeIt("FilterFixture/BM_FilterNative" + args + "/Threads:" + std::to_string(threads), iters,
                               rows, rows * 4, [&](Result&) -> arrow::Status {
                                 filter::FilterNative f{schema, batches};
                                 ARROW_RETURN_NOT_OK(f.Prepare());
                                 return f.Run().status();
                               }));
    if (want("BM_FilterGpu") && sys)
      results.push_back(TimeIt("FilterFixture/BM_FilterGpu" + args, iters, rows, rows * 4,
                               [&](Result& r) -> arrow::Status {
                                 filter::FilterGpu f{*sys, batches};
                                 ARROW_RETURN_NOT_OK(f.Prepare());
                                 ARROW_RETURN_NOT_OK(f.Run().status());
                                 AddTimers(r, f.Timers());
                                 return arrow::Status::OK();
                               }));
  }
  // ---- sum: SF batches of 2 Mi rows (aggr_benchmark.cc:132-150) ----
  if (want("BM_SumNative") || want("BM_SumGpu")) {
    RandomArrayGenerator rng(42);
    const int nb = sf, bs = 2 << 20;
    auto schema = vschema("v");
    auto batches = MakeRandomRecordBatches(rng, schema, nb, bs);
    gpu::PinnedBatches pinned(batches);
    const double rows = static_cast<double>(nb) * bs;
    const std::string args = "/Batches:" + std::to_string(nb) + "/BatchSize:" + std::to_string(bs);
    if (want("BM_SumNative"))
      results.push_back(TimeIt("AggregateFixture/BM_SumNative" + args + "/Threads:" + std::to_string(threads), iters,
                               rows, rows * 4, [&](Result&) -> arrow::Status {
                                 aggr::SumNative s{schema, batches};
                                 ARROW_RETURN_NOT_OK(s.Prepare());
                                 return s.Run().status();
                               }));
    if (want("BM_SumGpu") && sys)
      results.push_back(TimeIt("AggregateFixture/BM_SumGpu" + args, iters, rows, rows * 4,
                               [&](Result& r) -> arrow::Status {
                                 aggr::SumGpu s{*sys, batches};
                                 ARROW_RETURN_NOT_OK(s.Prepare());
                                 ARROW_RETURN_NOT_OK(s.Run().status());
                                 AddTimers(r, s.Timers());
                                 return arrow::Status::OK();
                               }));
  }
  // ---- take: SF batches of 4 Mi values, 512 Ki indices each (take_benchmark.cc:86-95,157-159) ----
  if (want("BM_TakeNative") || want("BM_TakeGpu")) {
    RandomArrayGenerator rng(42);
    const int nb = sf, bs = 4 << 20, ibs = bs >> 3;
    auto schema = vschema("v");
    auto batches = MakeRandomRecordBatches(rng, schema, nb, bs);
    auto md = arrow::key_value_metadata({{"min", "0"}, {"max", std::to_string(bs - 1)}});
    auto ischema = arrow::schema({arrow::field("i", arrow::uint32(), false, md)});
    auto indices = MakeRandomRecordBatches(rng, ischema, nb, ibs);
    gpu::PinnedBatches pinned(batches);
    pinned.Add(indices);
    const double rows = static_cast<double>(nb) * bs;  // the reference counts value rows (:54-56)
    const std::string args = "/Batches:" + std::to_string(nb) + "/BatchSize:" + std::to_string(bs);
    if (want("BM_TakeNative"))
      results.push_back(TimeIt("TakeFixture/BM_TakeNative" + args + "/Threads:" + std::to_string(threads), iters,
                               rows, rows * 4, [&](Result&) -> arrow::Status {
                                 take::TakeNative t{schema, batches, indices};
                                 return t.Run().status();
                               }));
    if (want("BM_TakeGpu") && sys)
      results.push_back(TimeIt("TakeFixture/BM_TakeGpu" + args, iters, rows, rows * 4,
                               [&](Result& r) -> arrow::Status {
                                 take::TakeGpu t{*sys, batches, indices};
                                 ARROW_RETURN_NOT_OK(t.Prepare());
                                 ARROW_RETURN_NOT_OK(t.Run().status());
                                 AddTimers(r, t.Timers());
                                 return arrow::Status::OK();
                               }));
  }
  // ---- join: SF batches of 2 Mi rows per side (join_benchmark.cc:83-100,168-176) ----
  if (want("BM_JoinNative") || want("BM_JoinGpu")) {
    RandomArrayGenerator rng(42);
    const int nb = sf, bs = 2 << 20;
    auto right = AddColumn("pk", MakeRandomRecordBatches(rng, vschema("x"), nb, bs), MakeIndexColumn(nb, bs).ValueOrDie());
    auto lefty = MakeRandomRecordBatches(rng, vschema("y"), nb, bs);
    auto left = AddColumn("fk", lefty, MakeForeignKeyColumn(rng, bs, nb, bs).ValueOrDie());
    gpu::PinnedBatches pinned(left);
    pinned.Add(right);
    const double items = 4.0 * nb * bs;  // rows x columns of both sides (join_benchmark.cc:114-125)
    const std::string args = "/Batches:" + std::to_string(nb) + "/BatchSize:" + std::to_string(bs);
    if (want("BM_JoinNative"))
      results.push_back(TimeIt("JoinFixture/BM_JoinNative" + args + "/Threads:" + std::to_string(threads), iters,
                               items, items * 4, [&](Result&) -> arrow::Status {
                                 join::JoinNative j{left[0]->schema(), right[0]->schema(), left, right};
                                 ARROW_RETURN_NOT_OK(j.Prepare());
                                 return j.Run().status();
                               }));
    if (want("BM_JoinGpu") && sys)
      results.push_back(TimeIt("JoinFixture/BM_JoinGpu" + args, iters, items, items * 4,
                               [&](Result& r) -> arrow::Status {
                                 join::JoinGpu j{*sys, left[0]->schema(), right[0]->schema(), left, right};
                                 ARROW_RETURN_NOT_OK(j.Prepare());
                                 ARROW_RETURN_NOT_OK(j.Run().status());
                                 AddTimers(r, j.Timers());
                                 return arrow::Status::OK();
                               }));
  }

  // ---- partition: SF batches of 2 Mi rows into 32 partitions (partition_benchmark.cc:66 is
  //      DISABLED_ in the reference: its DPU partition operator is broken, README.md:114-118) ----
  if (want("BM_PartitionGpu") && sys) {
    RandomArrayGenerator rng(42);
    const int nb = sf, bs = 2 << 20;
    auto batches = AddColumn("pk", MakeRandomRecordBatches(rng, vschema("x"), nb, bs), MakeIndexColumn(nb, bs).ValueOrDie());
    gpu::PinnedBatches pinned(batches);
    const double items = 2.0 * nb * bs;
    results.push_back(TimeIt("PartitionFixture/BM_PartitionGpu/Batches:" + std::to_string(nb) + "/BatchSize:" +
                                 std::to_string(bs) + "/Partitions:32",
                             iters, items, items * 4, [&](Result& r) -> arrow::Status {
                               partition::PartitionGpu p{*sys, batches[0]->schema(), batches, 32, "pk"};
                               ARROW_RETURN_NOT_OK(p.Prepare());
                               ARROW_RETURN_NOT_OK(p.Run().status());
                               AddTimers(r, p.Timers());
                               return arrow::Status::OK();
                             }));
  }

  // ---- gbench-shaped JSON (stdout, or --benchmark_out=FILE as scripts/run-*.sh use it) ----
  if (!out_path.empty() && !std::freopen(out_path.c_str(), "w", stdout)) return 4;
  std::printf("{\n  \"context\": {\"SF\": \"%d\", \"NR_GPUS\": \"%d\", \"host_threads\": \"%d\", \"arrow\": \"%s\"},\n",
              sf, sys ? sys->size() : 0, threads, ARROW_VERSION_STRING);
  std::printf("  \"benchmarks\": [\n");
  for (size_t i = 0; i < results.size(); ++i) {
    const Result& r = results[i];
    std::printf("    {\"name\": \"%s\", \"run_type\": \"iteration\", \"iterations\": %d, \"real_time\": %.6f, "
                "\"time_unit\": \"ms\", \"error_occurred\": %s",
                r.name.c_str(), r.iterations, r.real_ms, r.error ? "true" : "false");
    if (r.error) std::printf(", \"error_message\": \"%s\"", r.error_message.c_str());
    for (auto& kv : r.counters) std::printf(", \"%s\": %.6e", kv.first.c_str(), kv.second);
    std::printf("}%s\n", i + 1 < results.size() ? "," : "");
  }
  std::printf("  ]\n}\n");
  return 0;
}
