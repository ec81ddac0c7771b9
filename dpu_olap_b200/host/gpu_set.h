// gpu_set.h — the GPU-side counterpart of dpu::DpuSet (reference host/dpuext/dpuext.hpp:669-739):
// an RAII owner of one b2_ctx, plus the status mapping that replaces DPU_RETURN_NOT_OK
// (host/dpuext/status.h:7-12) and the timer::Timers shape the benchmarks read
// (host/timer/timer.h; filter_benchmark.cc:52-61).
#pragma once
#include <arrow/api.h>

#include <chrono>
#include <map>
#include <memory>
#include <string>

#include "b200olap.h"

#define B2_ARROW_RETURN_NOT_OK(ctx, expr)                                                    \
  do {                                                                                       \
    int _b2s = (expr);                                                                       \
    if (_b2s != B2_OK)                                                                       \
      return arrow::Status::UnknownError(b2_strerror(_b2s), ": ", b2_last_error(ctx));       \
  } while (0)

namespace gpu {

class GpuSet {
 public:
  static arrow::Result<std::shared_ptr<GpuSet>> allocate(int device = 0) {
    b2_ctx* ctx = nullptr;
    const int s = b2_ctx_create(device, &ctx);
    if (s != B2_OK) return arrow::Status::UnknownError("b2_ctx_create: ", b2_strerror(s));
    return std::shared_ptr<GpuSet>(new GpuSet(ctx));
  }
  ~GpuSet() { b2_ctx_destroy(ctx_); }
  GpuSet(const GpuSet&) = delete;
  GpuSet& operator=(const GpuSet&) = delete;
  b2_ctx* ctx() const { return ctx_; }

 private:
  explicit GpuSet(b2_ctx* ctx) : ctx_(ctx) {}
  b2_ctx* ctx_;
};

}  // namespace gpu

namespace timer {

// Same read interface as the reference's Timers (get() -> name -> Timer with Result()); the
// durations come from CUDA events inside the library (b2_timings), not from host clocks.
class Timer {
 public:
  explicit Timer(double ms = 0) : ns_(static_cast<int64_t>(ms * 1e6)) {}
  std::chrono::nanoseconds Result() const { return std::chrono::nanoseconds(ns_); }

 private:
  int64_t ns_;
};

class Timers {
 public:
  void Add(const b2_timings& t) {
    Acc("copy-to-dpu", t.copy_to_dev_ms);
    Acc("dpu-work", t.dev_work_ms);
    Acc("copy-from-dpu", t.copy_from_dev_ms);
    Acc("total", t.total_ms);
    h2d_bytes += t.h2d_bytes;
    d2h_bytes += t.d2h_bytes;
    kernel_launches += t.kernel_launches;
  }
  const std::map<std::string, std::shared_ptr<Timer>>& get() const { return timers_; }
  int64_t h2d_bytes = 0, d2h_bytes = 0, kernel_launches = 0;

 private:
  void Acc(const std::string& name, double ms) {
    double prev = 0;
    auto it = timers_.find(name);
    if (it != timers_.end()) prev = it->second->Result().count() / 1e6;
    timers_[name] = std::make_shared<Timer>(prev + ms);
  }
  std::map<std::string, std::shared_ptr<Timer>> timers_;
};

}  // namespace timer
