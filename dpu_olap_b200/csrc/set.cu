// set.cu — b2_set: the GPUs of one node behind ONE handle, driven by one host process.
//
// Replaces the reference's device set: dpu::DpuSet::allocate(nr_dpus) owns every DPU
// (host/dpuext/dpuext.hpp:704-739), the operators iterate over its ranks / DPUs ("batch i -> DPU i",
// filter_dpu.cc:127; groups of nr_dpus partitions, join_dpu.cc:254) and the join repartitions through
// the HOST between its phases (partitioner.cc:350-375). Here:
//   * filter / sum / take shard by contiguous batch ranges, one host thread per GPU driving that
//     GPU's own pipelined host entry point (no data-path collective; the sum adds one partial per GPU);
//   * the join has one exchange, fused into the routing kernel: every GPU counts its rows per
//     (destination GPU, coarse bucket), a plan kernel turns all GPUs' counts — read through peer
//     pointers — into destination addresses, the scatter kernel stores the (key, payload) pairs
//     straight into the peers' receive buffers over NVLink, and every GPU joins what it received.
//     Ordering between GPUs is CUDA events (cudaStreamWaitEvent across devices): no NCCL, no host
//     synchronisation between count and local join, no torch.
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"
#include "pending.h"

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

constexpr int kShuffleBits = 10;  // log2(#GPUs) destination bits + coarse bits: one radix pass

struct Member {
  b2_ctx* ctx = nullptr;
  // join state, grow-only
  uint64_t* recv[2] = {nullptr, nullptr};  // receive buffers (8-byte pairs): probe side, build side
  int64_t recv_cap = 0;                    // rows each holds
  char* tables = nullptr;                  // off[2][B+1] | off_ptrs[2][n] | recv_base[2][n] | addr[2][B] | seg[2][B+1] | info[2][3]
  cudaEvent_t ev_count = nullptr, ev_scatter = nullptr, ev_t0 = nullptr, ev_up = nullptr, ev_work = nullptr;
  int64_t* h_info = nullptr;  // pinned: info[2][3] | rows (or the member's b2_join_aggr: rows | sum_y | sum_x)
  // pending join result
  uint32_t* o_all = nullptr;  // fk | y | x, one allocation (adjacent output columns)
  int64_t o_cap = 0;
  uint64_t rows = 0;
};

struct TableLayout {
  size_t off, off_ptrs, recv_base, addr, seg, info, total;
};
TableLayout table_layout(int n) {
  const size_t B = (size_t)1 << kShuffleBits;
  TableLayout t;
  size_t o = 0;
  t.off = o;        o += b2_align_up(2 * (B + 1) * 8, 256);
  t.off_ptrs = o;   o += b2_align_up(2 * (size_t)n * 8, 256);
  t.recv_base = o;  o += b2_align_up(2 * (size_t)n * 8, 256);
  t.addr = o;       o += b2_align_up(2 * B * 8, 256);
  t.seg = o;        o += b2_align_up(2 * (B + 1) * 8, 256);
  t.info = o;       o += 256;
  t.total = o;
  return t;
}

int ensure_streams(b2_ctx* ctx) {
  if (!ctx->s_compute) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  if (!ctx->s_copy_in) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_in, cudaStreamNonBlocking));
  if (!ctx->s_copy_out) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_out, cudaStreamNonBlocking));
  return B2_OK;
}

// contiguous ranges, remainder to the low members (the reference deals batch i to DPU i)
void split(int64_t nbatches, int n, std::vector<int64_t>* first) {
  first->assign((size_t)n + 1, 0);
  const int64_t per = nbatches / n, rem = nbatches % n;
  for (int g = 0; g < n; ++g) (*first)[(size_t)g + 1] = (*first)[(size_t)g] + per + (g < rem ? 1 : 0);
}

void merge_timings(b2_timings* acc, const b2_timings& t) {  // members run concurrently: phases overlap
  acc->copy_to_dev_ms = std::max(acc->copy_to_dev_ms, t.copy_to_dev_ms);
  acc->dev_work_ms = std::max(acc->dev_work_ms, t.dev_work_ms);
  acc->copy_from_dev_ms = std::max(acc->copy_from_dev_ms, t.copy_from_dev_ms);
  acc->h2d_bytes += t.h2d_bytes;
  acc->d2h_bytes += t.d2h_bytes;
  acc->kernel_launches += t.kernel_launches;
}

}  // namespace

struct b2_set {
  std::vector<Member> m;
  std::string last_error;
  bool peer_ok = true;
  bool join_pending = false;
  bool filter_pending = false;
  std::vector<int64_t> filter_first;  // batch ranges of the pending filter result
  int n() const { return (int)m.size(); }
};

namespace {

int set_fail(b2_set* set, int status, const b2_ctx* from, const char* what) {
  set->last_error = std::string(what ? what : "") + ": " +
                    (from && !from->last_error.empty() ? from->last_error : std::string(b2_strerror(status)));
  return status;
}

// Runs fn(g) on one host thread per member and returns the first non-OK status.
template <typename F>
int for_each_member(b2_set* set, const char* what, F&& fn) {
  const int n = set->n();
  std::vector<int> rc((size_t)n, B2_OK);
  if (n == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g) th.emplace_back([&, g] { rc[(size_t)g] = fn(g); });
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < n; ++g)
    if (rc[(size_t)g] != B2_OK) return set_fail(set, rc[(size_t)g], set->m[(size_t)g].ctx, what);
  return B2_OK;
}

void free_join_result(Member& mb) {
  if (mb.o_all) {
    b2_device_scope sc(mb.ctx);
    b2_dev_free(mb.ctx, mb.o_all);
    mb.o_all = nullptr;
  }
  mb.o_cap = 0;
  mb.rows = 0;
}

// Per-member tables and events of the join (once), receive buffers of at least cap rows (grow-only).
int ensure_join_state(b2_set* set, int64_t cap) {
  const int n = set->n();
  const TableLayout T = table_layout(n);
  const size_t B = (size_t)1 << kShuffleBits;
  bool bases_changed = false;
  for (auto& mb : set->m) {
    b2_device_scope sc(mb.ctx);
    B2_RETURN_NOT_OK(ensure_streams(mb.ctx));
    if (!mb.tables) {
      B2_CUDA_OK(mb.ctx, cudaMalloc(&mb.tables, T.total));
      B2_CUDA_OK(mb.ctx, cudaMemset(mb.tables, 0, T.total));
      B2_CUDA_OK(mb.ctx, cudaEventCreateWithFlags(&mb.ev_count, cudaEventDisableTiming));
      B2_CUDA_OK(mb.ctx, cudaEventCreateWithFlags(&mb.ev_scatter, cudaEventDisableTiming));
      B2_CUDA_OK(mb.ctx, cudaEventCreate(&mb.ev_t0));
      B2_CUDA_OK(mb.ctx, cudaEventCreate(&mb.ev_up));
      B2_CUDA_OK(mb.ctx, cudaEventCreate(&mb.ev_work));
      B2_CUDA_OK(mb.ctx, cudaHostAlloc(&mb.h_info, 128, cudaHostAllocPortable));
      bases_changed = true;
    }
    if (mb.recv_cap < cap) {
      B2_CUDA_OK(mb.ctx, cudaDeviceSynchronize());
      for (int s = 0; s < 2; ++s) {
        if (mb.recv[s]) cudaFree(mb.recv[s]);
        mb.recv[s] = nullptr;
        B2_CUDA_OK(mb.ctx, cudaMalloc(&mb.recv[s], (size_t)cap * 8));
      }
      mb.recv_cap = cap;
      bases_changed = true;
    }
  }
  if (bases_changed) {
    // every member addresses every member's boundary table and receive buffers directly (peer access)
    std::vector<uint64_t> host(4 * (size_t)n);
    for (int s = 0; s < 2; ++s)
      for (int h = 0; h < n; ++h) {
        host[(size_t)s * n + h] = reinterpret_cast<uint64_t>(set->m[(size_t)h].tables + T.off + (size_t)s * (B + 1) * 8);
        host[2 * (size_t)n + (size_t)s * n + h] = reinterpret_cast<uint64_t>(set->m[(size_t)h].recv[s]);
      }
    for (auto& mb : set->m) {
      b2_device_scope sc(mb.ctx);
      B2_CUDA_OK(mb.ctx, cudaMemcpy(mb.tables + T.off_ptrs, host.data(), 2 * (size_t)n * 8, cudaMemcpyHostToDevice));
      B2_CUDA_OK(mb.ctx, cudaMemcpy(mb.tables + T.recv_base, host.data() + 2 * (size_t)n, 2 * (size_t)n * 8,
                                    cudaMemcpyHostToDevice));
    }
  }
  return B2_OK;
}

struct Inputs {  // one member's share of the join's input columns, on its device
  uint32_t *fk = nullptr, *y = nullptr, *pk = nullptr, *x = nullptr;
  int64_t nl = 0, nr = 0;
  void* ws[2] = {nullptr, nullptr};
  size_t ws_bytes[2] = {0, 0};
};

int upload_column(b2_ctx* ctx, uint32_t* d_col, const uint32_t* const* ptrs, const int64_t* lens, int64_t nbatches,
                  cudaStream_t s, int64_t* bytes) {
  int64_t off = 0, b = 0;
  while (b < nbatches) {  // host-adjacent batches travel as one copy
    int64_t e = b + 1, rows = lens[b];
    while (e < nbatches && ptrs[e] == ptrs[e - 1] + lens[e - 1]) rows += lens[e++];
    if (rows > 0) {
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_col + off, ptrs[b], (size_t)rows * 4, cudaMemcpyHostToDevice, s));
      *bytes += rows * 4;
    }
    off += rows;
    b = e;
  }
  return B2_OK;
}

}  // namespace

extern "C" {

int b2_set_create(const int* devices, int n, b2_set** out) {
  if (!out) return B2_ERR_INVALID;
  *out = nullptr;
  if (n < 1 || n > 64) return B2_ERR_INVALID;
  int prev = -1;
  cudaGetDevice(&prev);
  b2_set* set = new b2_set();
  int rc = B2_OK;
  for (int i = 0; i < n && rc == B2_OK; ++i) {
    const int dev = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (set->m[(size_t)j].ctx->device == dev) rc = B2_ERR_INVALID;  // a device appears once
    b2_ctx* ctx = nullptr;
    if (rc == B2_OK) rc = b2_ctx_create(dev, &ctx);
    if (rc == B2_OK) {
      Member mb;
      mb.ctx = ctx;
      set->m.push_back(mb);
    }
  }
  if (rc == B2_OK && n > 1) {
    // peer access between every pair: the join's scatter stores into peer memory, its plan kernel
    // reads the peers' boundary tables
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(set->m[(size_t)i].ctx->device);
      for (int j = 0; j < n; ++j) {
        if (i == j) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, set->m[(size_t)i].ctx->device, set->m[(size_t)j].ctx->device);
        if (!can) {
          set->peer_ok = false;
          continue;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(set->m[(size_t)j].ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) set->peer_ok = false;
        cudaGetLastError();
      }
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  if (rc != B2_OK) {
    for (auto& mb : set->m) b2_ctx_destroy(mb.ctx);
    delete set;
    return rc;
  }
  *out = set;
  return B2_OK;
}

int b2_set_destroy(b2_set* set) {
  if (!set) return B2_OK;
  for (auto& mb : set->m) {
    {
      b2_device_scope sc(mb.ctx);
      cudaDeviceSynchronize();
      free_join_result(mb);
      for (int s = 0; s < 2; ++s)
        if (mb.recv[s]) cudaFree(mb.recv[s]);
      if (mb.tables) cudaFree(mb.tables);
      for (cudaEvent_t e : {mb.ev_count, mb.ev_scatter, mb.ev_t0, mb.ev_up, mb.ev_work})
        if (e) cudaEventDestroy(e);
      if (mb.h_info) cudaFreeHost(mb.h_info);
    }
    b2_ctx_destroy(mb.ctx);
  }
  delete set;
  return B2_OK;
}

int b2_set_size(const b2_set* set) { return set ? set->n() : 0; }
b2_ctx* b2_set_ctx(b2_set* set, int i) { return (set && i >= 0 && i < set->n()) ? set->m[(size_t)i].ctx : nullptr; }
const char* b2_set_last_error(const b2_set* set) { return set ? set->last_error.c_str() : ""; }
int b2_set_peer_access(const b2_set* set) { return set && set->peer_ok ? 1 : 0; }
int64_t b2_set_launch_count(const b2_set* set) {
  int64_t n = 0;
  if (set)
    for (const auto& mb : set->m) n += mb.ctx->launches;
  return n;
}
int b2_set_set_inputs_pinned(b2_set* set, int on) {
  if (!set) return B2_ERR_INVALID;
  for (auto& mb : set->m) b2_ctx_set_inputs_pinned(mb.ctx, on);
  return B2_OK;
}

// ---- row-range sharded operators ------------------------------------------------------------------
int b2_set_sum_u32_host(b2_set* set, const uint32_t* const* batch_ptrs, const int64_t* batch_lens, int64_t nbatches,
                        uint64_t* sum, b2_timings* timings) {
  if (!set || !sum || nbatches < 0) return B2_ERR_INVALID;
  const auto t0 = Clock::now();
  const int n = set->n();
  std::vector<int64_t> first;
  split(nbatches, n, &first);
  std::vector<uint64_t> part((size_t)n, 0);
  std::vector<b2_timings> tm((size_t)n);
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_sum_u32_host", [&](int g) {
    const int64_t b0 = first[(size_t)g], nb = first[(size_t)g + 1] - b0;
    if (nb == 0) return (int)B2_OK;
    return b2_sum_u32_host(set->m[(size_t)g].ctx, batch_ptrs + b0, batch_lens + b0, nb, &part[(size_t)g], &tm[(size_t)g]);
  }));
  *sum = 0;
  b2_timings acc{};
  for (int g = 0; g < n; ++g) {  // one partial per device, added on the host (aggr_dpu.cc:82-84)
    *sum += part[(size_t)g];
    merge_timings(&acc, tm[(size_t)g]);
  }
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

int b2_set_filter_lt_u32_host(b2_set* set, const uint32_t* const* batch_ptrs, const int64_t* batch_lens,
                              int64_t nbatches, uint32_t threshold, int64_t* out_counts, uint64_t* total,
                              b2_timings* timings) {
  if (!set || nbatches < 0 || (nbatches > 0 && !out_counts)) return B2_ERR_INVALID;
  const auto t0 = Clock::now();
  const int n = set->n();
  split(nbatches, n, &set->filter_first);
  const std::vector<int64_t>& first = set->filter_first;
  std::vector<uint64_t> part((size_t)n, 0);
  std::vector<b2_timings> tm((size_t)n);
  set->filter_pending = false;
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_filter_lt_u32_host", [&](int g) {
    const int64_t b0 = first[(size_t)g], nb = first[(size_t)g + 1] - b0;
    return b2_filter_lt_u32_host(set->m[(size_t)g].ctx, batch_ptrs + b0, batch_lens + b0, nb, threshold,
                                 out_counts + b0, &part[(size_t)g], &tm[(size_t)g]);
  }));
  set->filter_pending = true;
  b2_timings acc{};
  uint64_t tot = 0;
  for (int g = 0; g < n; ++g) {
    tot += part[(size_t)g];
    merge_timings(&acc, tm[(size_t)g]);
  }
  if (total) *total = tot;
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

int b2_set_filter_fetch_host(b2_set* set, uint32_t* const* out_ptrs, int64_t nbatches, b2_timings* timings) {
  if (!set) return B2_ERR_INVALID;
  if (!set->filter_pending || set->filter_first.empty() || set->filter_first.back() != nbatches)
    return set_fail(set, B2_ERR_INVALID, nullptr, "b2_set_filter_fetch_host: no pending filter result of this shape");
  const auto t0 = Clock::now();
  const int n = set->n();
  const std::vector<int64_t>& first = set->filter_first;
  std::vector<b2_timings> tm((size_t)n);
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_filter_fetch_host", [&](int g) {
    const int64_t b0 = first[(size_t)g], nb = first[(size_t)g + 1] - b0;
    return b2_filter_fetch_host(set->m[(size_t)g].ctx, out_ptrs + b0, nb, &tm[(size_t)g]);
  }));
  b2_timings acc{};
  for (int g = 0; g < n; ++g) merge_timings(&acc, tm[(size_t)g]);
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

int b2_set_take_u32_host(b2_set* set, const uint32_t* const* value_ptrs, const int64_t* value_lens,
                         const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                         uint32_t* const* out_ptrs, b2_timings* timings) {
  if (!set || nbatches < 0) return B2_ERR_INVALID;
  const auto t0 = Clock::now();
  const int n = set->n();
  std::vector<int64_t> first;
  split(nbatches, n, &first);
  std::vector<b2_timings> tm((size_t)n);
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_take_u32_host", [&](int g) {
    const int64_t b0 = first[(size_t)g], nb = first[(size_t)g + 1] - b0;
    if (nb == 0) return (int)B2_OK;
    return b2_take_u32_host(set->m[(size_t)g].ctx, value_ptrs + b0, value_lens + b0, idx_ptrs + b0, idx_lens + b0, nb,
                            out_ptrs + b0, &tm[(size_t)g]);
  }));
  b2_timings acc{};
  for (int g = 0; g < n; ++g) merge_timings(&acc, tm[(size_t)g]);
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

// ---- join -----------------------------------------------------------------------------------------
// The sharded join; with `agg` the fused join -> aggregate pipeline: same exchange, every member's local
// join adds its output rows' payloads instead of materialising them, the host adds the members' sums.
struct SetJoinAgg {
  int filter_y;
  uint32_t y_threshold;
  b2_join_aggr* out;
};
static int set_join_core(b2_set* set, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                         const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches,
                         uint64_t* out_rows, b2_timings* timings, const SetJoinAgg* agg) {
  if (!set || !out_rows || nl_batches < 0 || nr_batches < 0) return B2_ERR_INVALID;
  const auto t0 = Clock::now();
  const int n = set->n();
  set->join_pending = false;
  for (auto& mb : set->m) free_join_result(mb);
  if (n == 1) {  // one device: the single-context join, whose result stays pending in that ctx
    if (agg) {
      const int rc = b2_join_aggr_u32_host(set->m[0].ctx, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches,
                                           agg->filter_y, agg->y_threshold, agg->out, timings);
      if (rc != B2_OK) return set_fail(set, rc, set->m[0].ctx, "b2_join_aggr_u32_host");
      *out_rows = agg->out->rows;
      return B2_OK;
    }
    const int rc = b2_join_u32_host(set->m[0].ctx, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches, out_rows,
                                    timings);
    if (rc != B2_OK) return set_fail(set, rc, set->m[0].ctx, "b2_join_u32_host");
    set->m[0].rows = *out_rows;
    set->join_pending = true;
    return B2_OK;
  }
  if ((n & (n - 1)) != 0)
    return set_fail(set, B2_ERR_UNSUPPORTED, nullptr, "the join shards over a power-of-two number of GPUs");
  if (!set->peer_ok)
    return set_fail(set, B2_ERR_UNSUPPORTED, nullptr, "the sharded join needs peer access between all GPUs of the set");
  int skip = 0;
  while ((1 << skip) < n) ++skip;
  const int seg_bits = kShuffleBits - skip;
  const size_t B = (size_t)1 << kShuffleBits;
  const TableLayout T = table_layout(n);
  int64_t nl_tot = 0, nr_tot = 0;
  for (int64_t b = 0; b < nl_batches; ++b) nl_tot += l_lens[b];
  for (int64_t b = 0; b < nr_batches; ++b) nr_tot += r_lens[b];
  // receive capacity: an even share plus 25 % and 64 Ki rows of slack for hash imbalance
  int64_t cap = std::max(nl_tot, nr_tot) / n;
  cap += cap / 4 + 65536;
  B2_RETURN_NOT_OK(ensure_join_state(set, cap));
  cap = set->m[0].recv_cap;
  for (auto& mb : set->m) cap = std::min(cap, mb.recv_cap);

  std::vector<int64_t> lfirst, rfirst;
  split(nl_batches, n, &lfirst);
  split(nr_batches, n, &rfirst);
  std::vector<Inputs> in((size_t)n);
  std::vector<b2_timings> tm((size_t)n);
  std::vector<int64_t> launches0((size_t)n);
  for (int g = 0; g < n; ++g) launches0[(size_t)g] = set->m[(size_t)g].ctx->launches;

  struct Cleanup {  // inputs and workspaces go back to the members' pools on every exit
    b2_set* set;
    std::vector<Inputs>* in;
    std::vector<void*> extra[64];
    ~Cleanup() {
      for (int g = 0; g < set->n(); ++g) {
        Member& mb = set->m[(size_t)g];
        b2_device_scope sc(mb.ctx);
        cudaStreamSynchronize(mb.ctx->s_compute);
        Inputs& I = (*in)[(size_t)g];
        for (void* p : {(void*)I.fk, (void*)I.y, (void*)I.pk, (void*)I.x, I.ws[0], I.ws[1]})
          if (p) b2_dev_free(mb.ctx, p);
        for (void* p : extra[g]) b2_dev_free(mb.ctx, p);
      }
    }
  } cleanup{set, &in, {}};

  // ---- phase 1, one host thread per GPU: upload this GPU's batch ranges, count rows per bucket ----
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_join_u32_host (upload + count)", [&](int g) {
    Member& mb = set->m[(size_t)g];
    b2_ctx* ctx = mb.ctx;
    b2_device_scope sc(ctx);
    b2_pending_free(ctx);
    Inputs& I = in[(size_t)g];
    const int64_t lb0 = lfirst[(size_t)g], lnb = lfirst[(size_t)g + 1] - lb0;
    const int64_t rb0 = rfirst[(size_t)g], rnb = rfirst[(size_t)g + 1] - rb0;
    for (int64_t b = 0; b < lnb; ++b) I.nl += l_lens[lb0 + b];
    for (int64_t b = 0; b < rnb; ++b) I.nr += r_lens[rb0 + b];
    cudaStream_t s = ctx->s_compute;
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&I.fk, (size_t)I.nl * 4));
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&I.y, (size_t)I.nl * 4));
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&I.pk, (size_t)I.nr * 4));
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&I.x, (size_t)I.nr * 4));
    I.ws_bytes[0] = b2_shuffle_p2p_ws_bytes(I.nl, kShuffleBits);
    I.ws_bytes[1] = b2_shuffle_p2p_ws_bytes(I.nr, kShuffleBits);
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &I.ws[0], I.ws_bytes[0]));
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &I.ws[1], I.ws_bytes[1]));
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_t0, s));
    b2_timings& t = tm[(size_t)g];
    // l_ptrs = [fk batches..., y batches...] of ALL batches; this member's range of each column
    B2_RETURN_NOT_OK(upload_column(ctx, I.fk, l_ptrs + lb0, l_lens + lb0, lnb, s, &t.h2d_bytes));
    B2_RETURN_NOT_OK(upload_column(ctx, I.y, l_ptrs + nl_batches + lb0, l_lens + lb0, lnb, s, &t.h2d_bytes));
    B2_RETURN_NOT_OK(upload_column(ctx, I.pk, r_ptrs + rb0, r_lens + rb0, rnb, s, &t.h2d_bytes));
    B2_RETURN_NOT_OK(upload_column(ctx, I.x, r_ptrs + nr_batches + rb0, r_lens + rb0, rnb, s, &t.h2d_bytes));
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_up, s));
    int64_t* off = reinterpret_cast<int64_t*>(mb.tables + T.off);
    // filter -> join -> aggregate: the predicate on the left payload is applied in front of the link
    B2_RETURN_NOT_OK(b2_shuffle_p2p_count_lt_dev(ctx, I.fk, I.y, I.nl, kShuffleBits, agg && agg->filter_y,
                                                 agg ? agg->y_threshold : 0u, off, I.ws[0], I.ws_bytes[0], s));
    B2_RETURN_NOT_OK(b2_shuffle_p2p_count_dev(ctx, I.pk, I.nr, kShuffleBits, off + (B + 1), I.ws[1], I.ws_bytes[1], s));
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_count, s));
    return (int)B2_OK;
  }));

  // ---- phase 2, enqueued from this thread: plan, scatter over NVLink, local join ----
  const int64_t nr_expected = std::max<int64_t>(nr_tot / n, 1);
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (int g = 0; g < n; ++g) {  // every boundary table is complete before anybody plans
      Member& mb = set->m[(size_t)g];
      b2_device_scope sc(mb.ctx);
      for (int h = 0; h < n; ++h)
        if (h != g) B2_CUDA_OK(mb.ctx, cudaStreamWaitEvent(mb.ctx->s_compute, set->m[(size_t)h].ev_count, 0));
    }
    for (int g = 0; g < n; ++g) {
      Member& mb = set->m[(size_t)g];
      b2_ctx* ctx = mb.ctx;
      b2_device_scope sc(ctx);
      cudaStream_t s = ctx->s_compute;
      Inputs& I = in[(size_t)g];
      const int64_t* const* off_ptrs = reinterpret_cast<const int64_t* const*>(mb.tables + T.off_ptrs);
      const uint64_t* recv_base = reinterpret_cast<const uint64_t*>(mb.tables + T.recv_base);
      uint64_t* addr = reinterpret_cast<uint64_t*>(mb.tables + T.addr);
      int64_t* seg = reinterpret_cast<int64_t*>(mb.tables + T.seg);
      int64_t* info = reinterpret_cast<int64_t*>(mb.tables + T.info);
      B2_RETURN_NOT_OK(b2_shuffle_p2p_plan_dev(ctx, off_ptrs, recv_base, g, n, kShuffleBits, cap, addr, seg, info,
                                               nullptr, s));
      B2_RETURN_NOT_OK(b2_shuffle_p2p_plan_dev(ctx, off_ptrs + n, recv_base + n, g, n, kShuffleBits, cap, addr + B,
                                               seg + (B + 1), info + 3, info + 2, s));
      B2_RETURN_NOT_OK(b2_shuffle_p2p_scatter_lt_dev(ctx, I.fk, I.y, I.nl, kShuffleBits, agg && agg->filter_y,
                                                     agg ? agg->y_threshold : 0u, addr, info + 2, I.ws[0],
                                                     I.ws_bytes[0], s));
      B2_RETURN_NOT_OK(b2_shuffle_p2p_scatter_dev(ctx, I.pk, I.x, I.nr, kShuffleBits, addr + B, info + 5, I.ws[1],
                                                  I.ws_bytes[1], s));
      B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_scatter, s));
    }
    for (int g = 0; g < n; ++g) {  // every peer's stores have landed before anybody joins
      Member& mb = set->m[(size_t)g];
      b2_device_scope sc(mb.ctx);
      for (int h = 0; h < n; ++h)
        if (h != g) B2_CUDA_OK(mb.ctx, cudaStreamWaitEvent(mb.ctx->s_compute, set->m[(size_t)h].ev_scatter, 0));
    }
    for (int g = 0; g < n; ++g) {
      Member& mb = set->m[(size_t)g];
      b2_ctx* ctx = mb.ctx;
      b2_device_scope sc(ctx);
      cudaStream_t s = ctx->s_compute;
      const int64_t* seg = reinterpret_cast<const int64_t*>(mb.tables + T.seg);
      const int64_t* info = reinterpret_cast<const int64_t*>(mb.tables + T.info);
      // PK-FK joins produce at most one row per received probe row; duplicate build keys can produce
      // more, in which case this member's local join is re-run below with the count it reported
      if (!agg && (!mb.o_all || mb.o_cap < cap)) {
        if (mb.o_all) b2_dev_free(ctx, mb.o_all);
        mb.o_all = nullptr;
        B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&mb.o_all, (size_t)cap * 12));
        mb.o_cap = cap;
      }
      void* jws = nullptr;
      const size_t jws_bytes = b2_join_seg_cap_ws_bytes(cap, cap, nr_expected, skip, seg_bits);
      B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &jws, jws_bytes));
      cleanup.extra[g].push_back(jws);
      uint64_t* d_rows = nullptr;
      B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_rows, 256));
      cleanup.extra[g].push_back(d_rows);
      if (agg)  // d_rows receives a b2_join_aggr: rows | sum_y | sum_x
        B2_RETURN_NOT_OK(b2_join_aggr_pairs_seg_cap_phased_dev(
            ctx, mb.recv[0], seg, cap, mb.recv[1], seg + (B + 1), cap, nr_expected, seg_bits, agg->filter_y,
            agg->y_threshold, reinterpret_cast<b2_join_aggr*>(d_rows), skip, info + 5, 7, jws, jws_bytes, s));
      else
        B2_RETURN_NOT_OK(b2_join_pairs_seg_cap_dev(ctx, mb.recv[0], seg, cap, mb.recv[1], seg + (B + 1), cap,
                                                   nr_expected, seg_bits, mb.o_all, mb.o_all + mb.o_cap,
                                                   mb.o_all + 2 * mb.o_cap, mb.o_cap, d_rows, skip, info + 5, jws,
                                                   jws_bytes, s));
      B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_work, s));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(mb.h_info, info, 48, cudaMemcpyDeviceToHost, s));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(mb.h_info + 6, d_rows, agg ? sizeof(b2_join_aggr) : 8, cudaMemcpyDeviceToHost, s));
    }
    bool overflow = false;
    int64_t need = 0;
    for (int g = 0; g < n; ++g) {
      Member& mb = set->m[(size_t)g];
      b2_device_scope sc(mb.ctx);
      B2_CUDA_OK(mb.ctx, cudaStreamSynchronize(mb.ctx->s_compute));
      overflow = overflow || mb.h_info[5] != 0;
      need = std::max(need, std::max(mb.h_info[1], mb.h_info[4]));
      mb.rows = (uint64_t)mb.h_info[6];
    }
    if (!overflow) break;
    // skewed keys: some GPU would have received more than its buffers hold. Nothing was stored (the
    // flag is the same on every GPU); the counts are still valid, so grow the buffers and go again.
    if (attempt == 1) return set_fail(set, B2_ERR_OVERFLOW, nullptr, "join: receive buffers overflowed twice");
    cap = need + need / 16 + 4096;
    B2_RETURN_NOT_OK(ensure_join_state(set, cap));
    for (auto& mb : set->m) free_join_result(mb);
  }
  // duplicate build keys: a member matched more rows than its output columns hold — re-run its local join
  for (int g = 0; g < n; ++g) {
    Member& mb = set->m[(size_t)g];
    if (mb.rows == ~0ull) return set_fail(set, B2_ERR_WORKSPACE, nullptr, "join: a partition buffer overflowed");
    if (agg || (int64_t)mb.rows <= mb.o_cap) continue;  // nothing is materialised: no capacity to outgrow
    b2_ctx* ctx = mb.ctx;
    b2_device_scope sc(ctx);
    cudaStream_t s = ctx->s_compute;
    const int64_t want = (int64_t)mb.rows;
    b2_dev_free(ctx, mb.o_all);
    mb.o_all = nullptr;
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&mb.o_all, (size_t)want * 12));
    mb.o_cap = want;
    const int64_t* seg = reinterpret_cast<const int64_t*>(mb.tables + T.seg);
    void* jws = nullptr;
    const size_t jws_bytes = b2_join_seg_cap_ws_bytes(cap, cap, nr_expected, skip, seg_bits);
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &jws, jws_bytes));
    cleanup.extra[g].push_back(jws);
    uint64_t* d_rows = nullptr;
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_rows, 256));
    cleanup.extra[g].push_back(d_rows);
    B2_RETURN_NOT_OK(b2_join_pairs_seg_cap_dev(ctx, mb.recv[0], seg, cap, mb.recv[1], seg + (B + 1), cap, nr_expected,
                                               seg_bits, mb.o_all, mb.o_all + mb.o_cap, mb.o_all + 2 * mb.o_cap, mb.o_cap,
                                               d_rows, skip, nullptr, jws, jws_bytes, s));
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_work, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(mb.h_info + 6, d_rows, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    mb.rows = (uint64_t)mb.h_info[6];
    if ((int64_t)mb.rows > mb.o_cap) return set_fail(set, B2_ERR_OVERFLOW, nullptr, "join: output larger than reported");
  }
  uint64_t total = 0;
  b2_timings acc{};
  for (int g = 0; g < n; ++g) {
    Member& mb = set->m[(size_t)g];
    total += mb.rows;
    float up = 0, work = 0;
    cudaEventElapsedTime(&up, mb.ev_t0, mb.ev_up);
    cudaEventElapsedTime(&work, mb.ev_up, mb.ev_work);
    b2_timings& t = tm[(size_t)g];
    t.copy_to_dev_ms = up;
    t.dev_work_ms = work;
    t.d2h_bytes = 56;
    t.kernel_launches = (int32_t)(mb.ctx->launches - launches0[(size_t)g]);
    merge_timings(&acc, t);
  }
  *out_rows = total;
  if (agg) {
    b2_join_aggr sum{};
    for (auto& mb : set->m) {  // h_info + 6: this member's b2_join_aggr
      sum.rows += (uint64_t)mb.h_info[6];
      sum.sum_y += (uint64_t)mb.h_info[7];
      sum.sum_x += (uint64_t)mb.h_info[8];
    }
    *agg->out = sum;
  } else {
    set->join_pending = true;
  }
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

int b2_set_join_u32_host(b2_set* set, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                         const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches,
                         uint64_t* out_rows, b2_timings* timings) {
  return set_join_core(set, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches, out_rows, timings, nullptr);
}

int b2_set_join_aggr_u32_host(b2_set* set, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                              const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches, int filter_y,
                              uint32_t y_threshold, b2_join_aggr* out, b2_timings* timings) {
  if (!out) return B2_ERR_INVALID;
  *out = b2_join_aggr{};
  const SetJoinAgg agg{filter_y, y_threshold, out};
  uint64_t rows = 0;
  return set_join_core(set, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches, &rows, timings, &agg);
}

int b2_set_join_fetch_host(b2_set* set, uint32_t* out_fk, uint32_t* out_y, uint32_t* out_x, int64_t capacity_rows,
                           b2_timings* timings) {
  if (!set) return B2_ERR_INVALID;
  if (!set->join_pending) return set_fail(set, B2_ERR_INVALID, nullptr, "b2_set_join_fetch_host: no pending join result");
  const auto t0 = Clock::now();
  const int n = set->n();
  if (n == 1) {
    const int rc = b2_join_fetch_host(set->m[0].ctx, out_fk, out_y, out_x, capacity_rows, timings);
    return rc == B2_OK ? rc : set_fail(set, rc, set->m[0].ctx, "b2_join_fetch_host");
  }
  std::vector<uint64_t> first((size_t)n + 1, 0);
  for (int g = 0; g < n; ++g) first[(size_t)g + 1] = first[(size_t)g] + set->m[(size_t)g].rows;
  if ((uint64_t)capacity_rows < first[(size_t)n])
    return set_fail(set, B2_ERR_OVERFLOW, nullptr, "b2_set_join_fetch_host: capacity_rows < result rows");
  if (first[(size_t)n] > 0 && !(out_fk && out_y && out_x)) return set_fail(set, B2_ERR_INVALID, nullptr, "null output column");
  std::vector<b2_timings> tm((size_t)n);
  // the result stays partition-major: GPU 0's rows, then GPU 1's, ... (row order is unspecified, as JoinDpu's)
  B2_RETURN_NOT_OK(for_each_member(set, "b2_set_join_fetch_host", [&](int g) {
    Member& mb = set->m[(size_t)g];
    b2_ctx* ctx = mb.ctx;
    if (mb.rows == 0) return (int)B2_OK;
    b2_device_scope sc(ctx);
    cudaStream_t s = ctx->s_copy_out;
    const size_t bytes = (size_t)mb.rows * 4;
    const uint64_t o = first[(size_t)g];
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_t0, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_fk + o, mb.o_all, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_y + o, mb.o_all + mb.o_cap, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_x + o, mb.o_all + 2 * mb.o_cap, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaEventRecord(mb.ev_up, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, mb.ev_t0, mb.ev_up);
    tm[(size_t)g].copy_from_dev_ms = ms;
    tm[(size_t)g].d2h_bytes = (int64_t)bytes * 3;
    return (int)B2_OK;
  }));
  b2_timings acc{};
  for (int g = 0; g < n; ++g) merge_timings(&acc, tm[(size_t)g]);
  acc.total_ms = ms_since(t0);
  if (timings) *timings = acc;
  return B2_OK;
}

}  // extern "C"
