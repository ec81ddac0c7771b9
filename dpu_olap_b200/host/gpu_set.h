// gpu_set.h — the GPU-side counterpart of dpu::DpuSet (reference host/dpuext/dpuext.hpp:669-739):
// an RAII owner of a b2_set (one b2_ctx per GPU), plus the status mapping that replaces DPU_RETURN_NOT_OK
// (host/dpuext/status.h:7-12) and the timer::Timers shape the benchmarks read
// (host/timer/timer.h; filter_benchmark.cc:52-61).
#pragma once
#include <arrow/api.h>

#include <chrono>
#include <map>
#include <memory>
#include <algorithm>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "b200olap.h"

#define B2_SET_RETURN_NOT_OK(set, expr)                                                      \
  do {                                                                                       \
    int _b2s = (expr);                                                                       \
    if (_b2s != B2_OK)                                                                       \
      return arrow::Status::UnknownError(b2_strerror(_b2s), ": ", b2_set_last_error(set));   \
  } while (0)

#define B2_ARROW_RETURN_NOT_OK(ctx, expr)                                                    \
  do {                                                                                       \
    int _b2s = (expr);                                                                       \
    if (_b2s != B2_OK)                                                                       \
      return arrow::Status::UnknownError(b2_strerror(_b2s), ": ", b2_last_error(ctx));       \
  } while (0)

namespace gpu {

// Page-locks the data buffers of record batches in place for as long as the object lives, so the
// library's host->device copies run at PCIe speed (zero-copy interop: the Arrow buffers themselves
// are the DMA source). Construct it where the benchmark builds its inputs (fixture SetUp, untimed).
class PinnedBatches {
 public:
  PinnedBatches() = default;
  explicit PinnedBatches(const arrow::RecordBatchVector& batches) { Add(batches); }
  ~PinnedBatches() {
    for (const void* p : regions_) b2_host_unregister(p);
  }
  PinnedBatches(const PinnedBatches&) = delete;
  PinnedBatches& operator=(const PinnedBatches&) = delete;
  void Add(const arrow::RecordBatchVector& batches) {
    for (const auto& b : batches)
      for (const auto& col : b->columns()) {
        const auto& buf = col->data()->buffers[1];
        if (buf && buf->size() > 0 && b2_host_register(buf->data(), static_cast<size_t>(buf->size())) == B2_OK)
          regions_.push_back(buf->data());
      }
  }
  size_t regions() const { return regions_.size(); }

 private:
  std::vector<const void*> regions_;
};

// Page-locked result memory. cudaHostAlloc costs ~0.6 ms per MiB, far more than the transfers it
// speeds up, so blocks are recycled: Acquire() hands out an arrow::Buffer whose last reference
// returns the block to the pool (results of one benchmark iteration are reused by the next).
class PinnedPool : public std::enable_shared_from_this<PinnedPool> {
 public:
  ~PinnedPool() {
    for (auto& blk : free_) b2_host_free_pinned(blk.first);
  }
  arrow::Result<std::shared_ptr<arrow::Buffer>> Acquire(int64_t bytes) {
    void* p = nullptr;
    int64_t cap = 0;
    {
      std::lock_guard<std::mutex> lock(mu_);
      for (size_t i = 0; i < free_.size(); ++i)
        if (free_[i].second >= bytes && (p == nullptr || free_[i].second < cap)) {
          p = free_[i].first;
          cap = free_[i].second;
        }
      if (p)
        for (size_t i = 0; i < free_.size(); ++i)
          if (free_[i].first == p) {
            free_.erase(free_.begin() + i);
            break;
          }
    }
    if (!p) {
      cap = std::max<int64_t>(bytes, 1 << 20);
      if (b2_host_alloc_pinned(static_cast<size_t>(cap), &p) != B2_OK)
        return arrow::Status::OutOfMemory("b2_host_alloc_pinned(", cap, ")");
    }
    std::weak_ptr<PinnedPool> weak = weak_from_this();
    auto* raw = new arrow::MutableBuffer(static_cast<uint8_t*>(p), bytes);
    return std::shared_ptr<arrow::Buffer>(raw, [weak, p, cap](arrow::Buffer* b) {
      delete b;
      if (auto pool = weak.lock()) {
        std::lock_guard<std::mutex> lock(pool->mu_);
        pool->free_.emplace_back(p, cap);
      } else {
        b2_host_free_pinned(p);
      }
    });
  }

 private:
  std::mutex mu_;
  std::vector<std::pair<void*, int64_t>> free_;
};

// dpu::DpuSet::allocate(nr_dpus) owns every DPU of the system (dpuext.hpp:704-739); GpuSet::allocate
// (nr_gpus) owns a b2_set: one b2_ctx per GPU of this process with peer access between them. The
// operator classes shard over all members (filter / sum / take by batch range, the join with the
// fused peer-store shuffle); ctx() is member 0 for the single-GPU entry points.
class GpuSet {
 public:
  static arrow::Result<std::shared_ptr<GpuSet>> allocate(int nr_gpus = 1, int first_device = 0) {
    if (nr_gpus < 1) return arrow::Status::Invalid("nr_gpus must be >= 1");
    std::vector<int> devs(static_cast<size_t>(nr_gpus));
    for (int i = 0; i < nr_gpus; ++i) devs[static_cast<size_t>(i)] = first_device + i;
    b2_set* set = nullptr;
    const int s = b2_set_create(devs.data(), nr_gpus, &set);
    if (s != B2_OK) return arrow::Status::UnknownError("b2_set_create: ", b2_strerror(s));
    return std::shared_ptr<GpuSet>(new GpuSet(set));
  }
  ~GpuSet() { b2_set_destroy(set_); }
  GpuSet(const GpuSet&) = delete;
  GpuSet& operator=(const GpuSet&) = delete;
  b2_set* set() const { return set_; }
  b2_ctx* ctx() const { return b2_set_ctx(set_, 0); }
  int size() const { return b2_set_size(set_); }
  PinnedPool& pinned() { return *pinned_; }
  // All record batches given to operators on this GpuSet are page-locked (gpu::PinnedBatches): groups
  // of batches are then uploaded by one gather kernel instead of one DMA per batch.
  void PromiseInputsPinned(bool on) { b2_set_set_inputs_pinned(set_, on ? 1 : 0); }

 private:
  explicit GpuSet(b2_set* set) : set_(set), pinned_(std::make_shared<PinnedPool>()) {}
  b2_set* set_;
  std::shared_ptr<PinnedPool> pinned_;
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
  // JoinDpu's timer names (join_dpu.cc:146-148); the probe kernel builds a partition's table and
  // probes it in one launch, so "probe" covers the reference's "build" + "probe"
  void Add(const b2_join_phases& p) {
    Acc("partitionKernel", p.partition_build_ms + p.partition_probe_ms);
    Acc("partitionKernel-build-side", p.partition_build_ms);
    Acc("partitionKernel-probe-side", p.partition_probe_ms);
    Acc("probe", p.probe_ms);
    Acc("take", p.take_ms);
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
