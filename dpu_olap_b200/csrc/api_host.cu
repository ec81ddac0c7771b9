// api_host.cu — the host-buffer entry points: what the reference's *Dpu operator classes did with
// dpu::DpuSet + arrow_copy_to_dpus / arrow_copy_from_dpus (host/dpuext/arrow_utils.cc:47-73,
// 147-266) now happens behind one C call per operator: upload the record batches' data buffers,
// run the kernels, hand the result back.
//
// The reference processes `nr_dpus` batches at a time (copy -> exec -> copy back per group,
// host/filter/filter_dpu.cc:128-157). Here a "group" is a chunk of consecutive batches (~64 MiB):
// chunk k+1 is uploaded on the copy stream while chunk k is processed on the compute stream and,
// where the result size is known up front (take), chunk k-1 is downloaded on a third stream.
// Consecutive batches that are adjacent in host memory are merged into one copy.
#include <chrono>
#include <vector>

#include "common.cuh"
#include "pending.h"

int b2_ctx_reserve_ws(b2_ctx* ctx, size_t bytes);  // gen.cu

void b2_pending_free(b2_ctx* ctx) {
  if (!ctx || !ctx->pending) return;
  for (void* p : ctx->pending->dev) b2_dev_free(ctx, p);
  delete ctx->pending;
  ctx->pending = nullptr;
}

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

constexpr int64_t kChunkBytes = 64ll << 20;

int ensure_streams(b2_ctx* ctx) {
  B2_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->s_compute) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  if (!ctx->s_copy_in) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_in, cudaStreamNonBlocking));
  if (!ctx->s_copy_out) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_out, cudaStreamNonBlocking));
  return B2_OK;
}

int dev_alloc(b2_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 256;
  return b2_dev_alloc(ctx, p, bytes);
}

// RAII list of scratch device allocations and events of one call.
struct Scratch {
  b2_ctx* owner = nullptr;
  std::vector<void*> dev;
  std::vector<cudaEvent_t> events;
  ~Scratch() {
    // An entry point may leave through an error path while kernels and DMAs are still in flight on
    // the ctx's streams: nothing they use is released (or recycled by the next call) before they
    // have drained. On the success path the streams are idle already and this costs microseconds.
    if (owner) {
      b2_device_scope sc(owner);
      for (cudaStream_t st : {owner->s_copy_in, owner->s_compute, owner->s_copy_out})
        if (st) cudaStreamSynchronize(st);
    }
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    for (void* p : dev) b2_dev_free(owner, p);  // back to the ctx's recycling pool
  }
  int alloc(b2_ctx* ctx, void** p, size_t bytes) {
    owner = ctx;
    B2_RETURN_NOT_OK(dev_alloc(ctx, p, bytes));
    dev.push_back(*p);
    return B2_OK;
  }
  int event(b2_ctx* ctx, cudaEvent_t* e, bool timing) {
    B2_CUDA_OK(ctx, cudaEventCreateWithFlags(e, timing ? cudaEventDefault : cudaEventDisableTiming));
    events.push_back(*e);
    return B2_OK;
  }
};

struct Layout {
  std::vector<int64_t> off;  // nbatches+1 row offsets of the packed device column
  bool uniform = true;
  int64_t batch_len = 0;
  int64_t rows() const { return off.back(); }
};

int make_layout(b2_ctx* ctx, const uint32_t* const* ptrs, const int64_t* lens, int64_t nbatches,
                Layout* L) {
  B2_REQUIRE(ctx, nbatches >= 0, "negative batch count");
  B2_REQUIRE(ctx, nbatches == 0 || (ptrs && lens), "null batch table");
  L->off.assign((size_t)nbatches + 1, 0);
  L->batch_len = nbatches > 0 ? lens[0] : 0;
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, lens[b] >= 0, "negative batch length");
    B2_REQUIRE(ctx, lens[b] == 0 || ptrs[b] != nullptr, "null batch pointer");
    if (lens[b] != L->batch_len) L->uniform = false;
    L->off[(size_t)b + 1] = L->off[(size_t)b] + lens[b];
  }
  return B2_OK;
}

// Chunk boundaries: consecutive batches, about kChunkBytes each, at least one batch per chunk.
std::vector<int64_t> make_chunks(const Layout& L, int64_t nbatches) {
  std::vector<int64_t> c{0};
  int64_t acc = 0;
  for (int64_t b = 0; b < nbatches; ++b) {
    const int64_t bytes = (L.off[(size_t)b + 1] - L.off[(size_t)b]) * 4;
    if (acc > 0 && acc + bytes > kChunkBytes) {
      c.push_back(b);
      acc = 0;
    }
    acc += bytes;
  }
  if (nbatches > 0) c.push_back(nbatches);
  return c;
}

// ---- gather upload: one kernel reads a group of page-locked host batches over PCIe ------------
constexpr int64_t kGatherPiece = 16384;  // rows per CTA (64 KB)
struct GatherEntry {
  const uint32_t* src;  // host pointer, device-accessible (mapped pinned memory)
  uint32_t* dst;
  int64_t rows;
};

__global__ void __launch_bounds__(256)
gather_host_batches_kernel(const GatherEntry* __restrict__ tbl) {
  const GatherEntry e = tbl[blockIdx.x];  // the table itself lives in pinned host memory
  const uint32_t tid = threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(e.src) | reinterpret_cast<uintptr_t>(e.dst)) & 15) == 0) {
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(e.src);
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(e.dst);
    const int64_t n4 = e.rows >> 2;
    for (int64_t base = 0; base < n4; base += 256 * 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = base + u * 256 + tid;
        if (i < n4) v[u] = ld_stream_v4(s4 + i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = base + u * 256 + tid;
        if (i < n4) d4[i] = v[u];
      }
    }
    for (int64_t i = (n4 << 2) + tid; i < e.rows; i += 256) e.dst[i] = ld_stream_u32(e.src + i);
  } else {
    for (int64_t i = tid; i < e.rows; i += 256) e.dst[i] = ld_stream_u32(e.src + i);
  }
}

// Reserves room for `entries` more gather entries in the ctx's pinned table (call before the first
// upload of a *_host call, when nothing is in flight). Returns false if the gather path is off.
bool gather_begin(b2_ctx* ctx, int64_t nbatches, int64_t total_rows) {
  ctx->gather_used = 0;
  if (!ctx->inputs_pinned) return false;
  const size_t need = (size_t)(nbatches + total_rows / kGatherPiece + 16) * sizeof(GatherEntry);
  if (ctx->gather_bytes < need) {
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    ctx->h_gather = nullptr;
    ctx->gather_bytes = 0;
    const size_t want = b2_align_up(need, 1 << 16);
    if (cudaHostAlloc(&ctx->h_gather, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    ctx->gather_bytes = want;
  }
  return true;
}

// Upload batches [b0,b1) to their packed positions, merging host-adjacent batches. With
// b2_ctx_set_inputs_pinned the group goes up through ONE gather kernel (no per-batch DMA set-up).
int upload(b2_ctx* ctx, uint32_t* d_col, const Layout& L, const uint32_t* const* ptrs, int64_t b0,
           int64_t b1, cudaStream_t s, int64_t* bytes) {
  // count the copies the DMA path would need
  int64_t ncopies = 0, rows_total = L.off[(size_t)b1] - L.off[(size_t)b0];
  for (int64_t b = b0; b < b1;) {
    int64_t e = b + 1;
    while (e < b1 && ptrs[e] == ptrs[e - 1] + (L.off[(size_t)e] - L.off[(size_t)e - 1])) ++e;
    if (L.off[(size_t)e] > L.off[(size_t)b]) ++ncopies;
    b = e;
  }
  const size_t max_entries = (size_t)(ncopies + rows_total / kGatherPiece + 1);
  if (ctx->inputs_pinned && ncopies >= 4 && ctx->h_gather &&
      ctx->gather_used + max_entries * sizeof(GatherEntry) <= ctx->gather_bytes) {
    GatherEntry* tbl = reinterpret_cast<GatherEntry*>(static_cast<char*>(ctx->h_gather) + ctx->gather_used);
    int64_t n = 0;
    for (int64_t b = b0; b < b1;) {
      int64_t e = b + 1;
      while (e < b1 && ptrs[e] == ptrs[e - 1] + (L.off[(size_t)e] - L.off[(size_t)e - 1])) ++e;
      const int64_t rows = L.off[(size_t)e] - L.off[(size_t)b];
      for (int64_t r = 0; r < rows; r += kGatherPiece) {
        tbl[n].src = ptrs[b] + r;
        tbl[n].dst = d_col + L.off[(size_t)b] + r;
        tbl[n].rows = std::min(kGatherPiece, rows - r);
        ++n;
      }
      b = e;
    }
    ctx->gather_used += (size_t)n * sizeof(GatherEntry);
    if (n > 0) {
      gather_host_batches_kernel<<<(unsigned)n, 256, 0, s>>>(tbl);  // mapped: host pointer == device pointer
      B2_LAUNCH_CHECK(ctx, "gather_host_batches_kernel");
      *bytes += rows_total * 4;
    }
    return B2_OK;
  }
  int64_t b = b0;
  while (b < b1) {
    int64_t e = b + 1;
    while (e < b1 && ptrs[e] == ptrs[e - 1] + (L.off[(size_t)e] - L.off[(size_t)e - 1])) ++e;
    const int64_t rows = L.off[(size_t)e] - L.off[(size_t)b];
    if (rows > 0) {
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_col + L.off[(size_t)b], ptrs[b], (size_t)rows * 4,
                                      cudaMemcpyHostToDevice, s));
      *bytes += rows * 4;
    }
    b = e;
  }
  return B2_OK;
}

// Download packed rows [L.off[b0], L.off[b1]) to per-batch host pointers (merging adjacent ones).
int download(b2_ctx* ctx, const uint32_t* d_col, const Layout& L, uint32_t* const* ptrs, int64_t b0,
             int64_t b1, cudaStream_t s, int64_t* bytes) {
  int64_t b = b0;
  while (b < b1) {
    int64_t e = b + 1;
    while (e < b1 && ptrs[e] == ptrs[e - 1] + (L.off[(size_t)e] - L.off[(size_t)e - 1])) ++e;
    const int64_t rows = L.off[(size_t)e] - L.off[(size_t)b];
    if (rows > 0) {
      B2_CUDA_OK(ctx, cudaMemcpyAsync(ptrs[b], d_col + L.off[(size_t)b], (size_t)rows * 4,
                                      cudaMemcpyDeviceToHost, s));
      *bytes += rows * 4;
    }
    b = e;
  }
  return B2_OK;
}

struct PhaseTimer {  // sums CUDA-event intervals recorded on one stream
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
  double total_ms() const {
    double t = 0;
    for (auto& sp : spans) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, sp.first, sp.second) == cudaSuccess) t += ms;
    }
    return t;
  }
};

}  // namespace

extern "C" {

// ---- Sum ----------------------------------------------------------------------------------
int b2_sum_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const int64_t* batch_lens,
                    int64_t nbatches, uint64_t* sum, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, sum != nullptr, "sum is null");
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout L;
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  const std::vector<int64_t> chunks = make_chunks(L, nbatches);
  const size_t nchunks = chunks.size() - 1;
  gather_begin(ctx, nbatches, L.rows());
  *sum = 0;
  b2_timings tm{};
  if (nchunks > 0 && L.rows() > 0) {
    Scratch sc;
    // double-buffered chunk slots + one partial per chunk
    int64_t slot_rows = 0;
    for (size_t k = 0; k < nchunks; ++k)
      slot_rows = std::max(slot_rows, L.off[(size_t)chunks[k + 1]] - L.off[(size_t)chunks[k]]);
    const int nslots = nchunks > 1 ? 3 : 1;
    uint32_t* d_slots = nullptr;
    uint64_t* d_part = nullptr;
    // grow-only buffers owned by the ctx: no cudaMalloc / cudaFree per call
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 0, (size_t)slot_rows * 4 * nslots, (void**)&d_slots));
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 3, nchunks * 8, (void**)&d_part));
    std::vector<cudaEvent_t> up_done(nchunks), k_begin(nchunks), k_end(nchunks);
    cudaEvent_t c_begin, c_end;
    B2_RETURN_NOT_OK(sc.event(ctx, &c_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &c_end, true));
    PhaseTimer work;
    B2_CUDA_OK(ctx, cudaEventRecord(c_begin, ctx->s_copy_in));
    for (size_t k = 0; k < nchunks; ++k) {
      B2_RETURN_NOT_OK(sc.event(ctx, &up_done[k], false));
      B2_RETURN_NOT_OK(sc.event(ctx, &k_begin[k], true));
      B2_RETURN_NOT_OK(sc.event(ctx, &k_end[k], true));
      uint32_t* slot = d_slots + (size_t)(k % nslots) * slot_rows;
      // the slot is free once the kernel that used it (chunk k - nslots) has finished
      if (k >= (size_t)nslots) B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_copy_in, k_end[k - nslots], 0));
      // upload with a layout rebased to the slot
      const int64_t row0 = L.off[(size_t)chunks[k]];
      B2_RETURN_NOT_OK(upload(ctx, slot - row0, L, batch_ptrs, chunks[k], chunks[k + 1],
                              ctx->s_copy_in, &tm.h2d_bytes));
      B2_CUDA_OK(ctx, cudaEventRecord(up_done[k], ctx->s_copy_in));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_compute, up_done[k], 0));
      B2_CUDA_OK(ctx, cudaEventRecord(k_begin[k], ctx->s_compute));
      B2_RETURN_NOT_OK(b2_sum_u32_dev(ctx, slot, L.off[(size_t)chunks[k + 1]] - row0, d_part + k,
                                      ctx->s_compute));
      B2_CUDA_OK(ctx, cudaEventRecord(k_end[k], ctx->s_compute));
      work.spans.push_back({k_begin[k], k_end[k]});
    }
    B2_CUDA_OK(ctx, cudaEventRecord(c_end, ctx->s_copy_in));
    std::vector<uint64_t> parts(nchunks);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(parts.data(), d_part, nchunks * 8, cudaMemcpyDeviceToHost,
                                    ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_in));
    for (uint64_t p : parts) *sum += p;  // host adds the partials, as aggr_dpu.cc:82-84
    float ms = 0;
    cudaEventElapsedTime(&ms, c_begin, c_end);
    tm.copy_to_dev_ms = ms;
    tm.dev_work_ms = work.total_ms();
    tm.d2h_bytes = (int64_t)nchunks * 8;
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

// ---- Filter -------------------------------------------------------------------------------
int b2_filter_lt_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                          const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                          int64_t* out_counts, uint64_t* total, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nbatches == 0 || out_counts != nullptr, "out_counts is null");
  Layout L;
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  const std::vector<int64_t> chunks = make_chunks(L, nbatches);
  gather_begin(ctx, nbatches, L.rows());
  const size_t nchunks = chunks.size() - 1;
  b2_timings tm{};
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kFilter;
  pend->batch_end.assign((size_t)nbatches, 0);
  ctx->pending = pend;
  if (total) *total = 0;
  if (nbatches > 0) {
    Scratch sc;
    const int64_t n = L.rows();
    uint32_t *d_in = nullptr, *d_out = nullptr;
    int64_t *d_end = nullptr, *d_off = nullptr, *d_carry = nullptr;
    void* d_ws = nullptr;
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_in, (size_t)n * 4));
    B2_RETURN_NOT_OK(dev_alloc(ctx, (void**)&d_out, (size_t)n * 4));
    pend->dev.push_back(d_out);
    pend->d_out = d_out;
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_end, (size_t)nbatches * 8));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_carry, (nchunks + 1) * 8));
    B2_CUDA_OK(ctx, cudaMemsetAsync(d_carry, 0, 8, ctx->s_compute));
    size_t ws_bytes = 0;
    for (size_t k = 0; k < nchunks; ++k) {
      const int64_t nb = chunks[k + 1] - chunks[k];
      const size_t w = L.uniform ? b2_filter_ws_bytes(nb, L.batch_len)
                                 : b2_filter_ragged_ws_bytes(&L.off[(size_t)chunks[k]], nb);
      ws_bytes = std::max(ws_bytes, b2_align_up(w, 256));
    }
    B2_RETURN_NOT_OK(sc.alloc(ctx, &d_ws, ws_bytes * 2));  // alternate: memset of k+1 vs kernel k
    if (!L.uniform) {
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_off, (size_t)(nbatches + 1) * 8));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_off, L.off.data(), (size_t)(nbatches + 1) * 8,
                                      cudaMemcpyHostToDevice, ctx->s_compute));
    }
    cudaEvent_t c_begin, c_end;
    B2_RETURN_NOT_OK(sc.event(ctx, &c_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &c_end, true));
    PhaseTimer work;
    B2_CUDA_OK(ctx, cudaEventRecord(c_begin, ctx->s_copy_in));
    for (size_t k = 0; k < nchunks; ++k) {
      cudaEvent_t up_done, kb, ke;
      B2_RETURN_NOT_OK(sc.event(ctx, &up_done, false));
      B2_RETURN_NOT_OK(sc.event(ctx, &kb, true));
      B2_RETURN_NOT_OK(sc.event(ctx, &ke, true));
      B2_RETURN_NOT_OK(upload(ctx, d_in, L, batch_ptrs, chunks[k], chunks[k + 1], ctx->s_copy_in,
                              &tm.h2d_bytes));
      B2_CUDA_OK(ctx, cudaEventRecord(up_done, ctx->s_copy_in));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_compute, up_done, 0));
      B2_CUDA_OK(ctx, cudaEventRecord(kb, ctx->s_compute));
      const int64_t b0 = chunks[k], nb = chunks[k + 1] - chunks[k];
      void* ws_k = static_cast<char*>(d_ws) + (k & 1) * ws_bytes;
      int rc;
      // chunk k continues the output where chunk k-1 stopped: its running count is seeded from
      // d_carry[k] (device side, no host round trip) and it leaves its own total in d_carry[k+1]
      if (L.uniform) {
        rc = b2_filter_lt_u32_dev(ctx, d_in + L.off[(size_t)b0], nb, L.batch_len, threshold, d_out,
                                  d_end + b0, d_carry + k + 1, d_carry + k, ws_k, ws_bytes,
                                  ctx->s_compute);
      } else {
        rc = b2_filter_lt_u32_ragged_dev(ctx, d_in, &L.off[(size_t)b0], d_off + b0, nb, threshold,
                                         d_out, d_end + b0, d_carry + k + 1, d_carry + k, ws_k,
                                         ws_bytes, ctx->s_compute);
      }
      B2_RETURN_NOT_OK(rc);
      B2_CUDA_OK(ctx, cudaEventRecord(ke, ctx->s_compute));
      work.spans.push_back({kb, ke});
    }
    B2_CUDA_OK(ctx, cudaEventRecord(c_end, ctx->s_copy_in));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(pend->batch_end.data(), d_end, (size_t)nbatches * 8,
                                    cudaMemcpyDeviceToHost, ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_in));
    int64_t prev = 0;
    for (int64_t b = 0; b < nbatches; ++b) {
      out_counts[b] = pend->batch_end[(size_t)b] - prev;
      prev = pend->batch_end[(size_t)b];
    }
    if (total) *total = (uint64_t)prev;
    float ms = 0;
    cudaEventElapsedTime(&ms, c_begin, c_end);
    tm.copy_to_dev_ms = ms;
    tm.dev_work_ms = work.total_ms();
    tm.d2h_bytes = nbatches * 8;
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_filter_fetch_host(b2_ctx* ctx, uint32_t* const* out_ptrs, int64_t nbatches,
                         b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  b2_pending* pend = ctx->pending;
  if (!pend || pend->kind != b2_pending::kFilter)
    return b2_set_error(ctx, B2_ERR_INVALID, "b2_filter_fetch_host", "no pending filter result");
  B2_REQUIRE(ctx, nbatches == (int64_t)pend->batch_end.size(), "batch count differs from the run");
  B2_REQUIRE(ctx, nbatches == 0 || out_ptrs != nullptr, "out_ptrs is null");
  b2_timings tm{};
  if (nbatches > 0) {
    Layout R;  // layout of the compacted result: batch b = [end[b-1], end[b])
    R.off.assign((size_t)nbatches + 1, 0);
    for (int64_t b = 0; b < nbatches; ++b) R.off[(size_t)b + 1] = pend->batch_end[(size_t)b];
    for (int64_t b = 0; b < nbatches; ++b)
      B2_REQUIRE(ctx, R.off[(size_t)b + 1] == R.off[(size_t)b] || out_ptrs[b] != nullptr,
                 "null output pointer for a non-empty chunk");
    cudaEvent_t e0, e1;
    Scratch sc;
    B2_RETURN_NOT_OK(sc.event(ctx, &e0, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &e1, true));
    B2_CUDA_OK(ctx, cudaEventRecord(e0, ctx->s_copy_out));
    B2_RETURN_NOT_OK(download(ctx, pend->d_out, R, out_ptrs, 0, nbatches, ctx->s_copy_out, &tm.d2h_bytes));
    B2_CUDA_OK(ctx, cudaEventRecord(e1, ctx->s_copy_out));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_out));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    tm.copy_from_dev_ms = ms;
  }
  tm.total_ms = ms_since(t0);
  if (timings) *timings = tm;
  return B2_OK;
}

// Streaming variant: upload chunk k+1 | filter chunk k | download the output of chunk k-1 all
// overlap (PCIe is full duplex), and nothing is allocated per call. The only host round trip per
// chunk is the 8-byte running count that sizes its download.
int b2_filter_lt_u32_host_into(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                               const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                               uint32_t* out, int64_t out_capacity, int64_t* out_counts,
                               uint64_t* total, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nbatches == 0 || out_counts != nullptr, "out_counts is null");
  B2_REQUIRE(ctx, out_capacity >= 0 && (out_capacity == 0 || out != nullptr), "bad output buffer");
  Layout L;
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  const std::vector<int64_t> chunks = make_chunks(L, nbatches);
  gather_begin(ctx, nbatches, L.rows());
  const size_t nchunks = chunks.size() - 1;
  b2_timings tm{};
  if (total) *total = 0;
  int status = B2_OK;
  if (nbatches > 0) {
    Scratch sc;  // events only
    const int64_t n = L.rows();
    constexpr int kSlots = 3;
    int64_t slot_rows = 0;
    size_t ws_bytes = 0;
    for (size_t k = 0; k < nchunks; ++k) {
      const int64_t nb = chunks[k + 1] - chunks[k];
      slot_rows = std::max(slot_rows, L.off[(size_t)chunks[k + 1]] - L.off[(size_t)chunks[k]]);
      const size_t w = L.uniform ? b2_filter_ws_bytes(nb, L.batch_len)
                                 : b2_filter_ragged_ws_bytes(&L.off[(size_t)chunks[k]], nb);
      ws_bytes = std::max(ws_bytes, b2_align_up(w, 256));
    }
    slot_rows = (int64_t)b2_align_up((size_t)slot_rows, 64);
    // cache slot 0: input ring | 1: compacted output | 2: small tables + workspaces
    void *p_in = nullptr, *p_out = nullptr, *p_misc = nullptr;
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 0, (size_t)slot_rows * 4 * kSlots, &p_in));
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 1, (size_t)std::max<int64_t>(n, 1) * 4, &p_out));
    const size_t end_bytes = b2_align_up((size_t)nbatches * 8, 256);
    const size_t carry_bytes = b2_align_up((nchunks + 1) * 8, 256);
    const size_t off_bytes = L.uniform ? 0 : b2_align_up((size_t)kSlots * (size_t)(nbatches + 1) * 8, 256);
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 2, end_bytes + carry_bytes + off_bytes + ws_bytes * kSlots, &p_misc));
    uint32_t* d_in = static_cast<uint32_t*>(p_in);
    uint32_t* d_out = static_cast<uint32_t*>(p_out);
    char* misc = static_cast<char*>(p_misc);
    int64_t* d_end = reinterpret_cast<int64_t*>(misc);
    int64_t* d_carry = reinterpret_cast<int64_t*>(misc + end_bytes);
    int64_t* d_off = reinterpret_cast<int64_t*>(misc + end_bytes + carry_bytes);
    char* d_ws = misc + end_bytes + carry_bytes + off_bytes;
    // pinned mirror of the running counts (one per chunk) so the host can size the downloads
    if (ctx->pinned_bytes < (nchunks + 1) * 8) {
      if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
      ctx->h_pinned = nullptr;
      ctx->pinned_bytes = 0;
      const size_t want = b2_align_up((nchunks + 1) * 8, 4096);
      B2_CUDA_OK(ctx, cudaMallocHost(&ctx->h_pinned, want));
      ctx->pinned_bytes = want;
    }
    volatile int64_t* h_carry = static_cast<volatile int64_t*>(ctx->h_pinned);
    h_carry[0] = 0;
    B2_CUDA_OK(ctx, cudaMemsetAsync(d_carry, 0, 8, ctx->s_compute));
    std::vector<int64_t> rel;  // ragged: chunk-relative offsets
    std::vector<cudaEvent_t> k_end(nchunks), c_ready(nchunks);
    cudaEvent_t c_begin, c_end, o_begin, o_end;
    B2_RETURN_NOT_OK(sc.event(ctx, &c_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &c_end, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &o_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &o_end, true));
    PhaseTimer work;
    B2_CUDA_OK(ctx, cudaEventRecord(c_begin, ctx->s_copy_in));
    B2_CUDA_OK(ctx, cudaEventRecord(o_begin, ctx->s_copy_out));
    size_t drained = 0;  // chunks whose output download has been issued
    auto drain = [&](size_t upto) -> int {  // issue the downloads of chunks [drained, upto)
      for (; drained < upto; ++drained) {
        B2_CUDA_OK(ctx, cudaEventSynchronize(c_ready[drained]));
        const int64_t lo = h_carry[drained], hi = h_carry[drained + 1];
        if (hi > out_capacity) { status = B2_ERR_OVERFLOW; continue; }
        if (hi > lo) {
          B2_CUDA_OK(ctx, cudaMemcpyAsync(out + lo, d_out + lo, (size_t)(hi - lo) * 4,
                                          cudaMemcpyDeviceToHost, ctx->s_copy_out));
          tm.d2h_bytes += (hi - lo) * 4;
        }
      }
      return B2_OK;
    };
    for (size_t k = 0; k < nchunks; ++k) {
      cudaEvent_t up_done, kb;
      B2_RETURN_NOT_OK(sc.event(ctx, &up_done, false));
      B2_RETURN_NOT_OK(sc.event(ctx, &kb, true));
      B2_RETURN_NOT_OK(sc.event(ctx, &k_end[k], true));
      B2_RETURN_NOT_OK(sc.event(ctx, &c_ready[k], false));
      const int64_t b0 = chunks[k], b1 = chunks[k + 1], nb = b1 - b0;
      const int64_t row0 = L.off[(size_t)b0];
      uint32_t* slot = d_in + (size_t)(k % kSlots) * slot_rows;
      if (k >= (size_t)kSlots) B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_copy_in, k_end[k - kSlots], 0));
      B2_RETURN_NOT_OK(upload(ctx, slot - row0, L, batch_ptrs, b0, b1, ctx->s_copy_in, &tm.h2d_bytes));
      int64_t* d_off_k = nullptr;
      if (!L.uniform) {
        rel.resize((size_t)nb + 1);
        for (int64_t b = 0; b <= nb; ++b) rel[(size_t)b] = L.off[(size_t)(b0 + b)] - row0;
        d_off_k = d_off + (size_t)(k % kSlots) * (size_t)(nbatches + 1);
        B2_CUDA_OK(ctx, cudaMemcpyAsync(d_off_k, rel.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice,
                                        ctx->s_copy_in));  // pageable: staged before the call returns
      }
      B2_CUDA_OK(ctx, cudaEventRecord(up_done, ctx->s_copy_in));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_compute, up_done, 0));
      B2_CUDA_OK(ctx, cudaEventRecord(kb, ctx->s_compute));
      void* ws_k = d_ws + (k % kSlots) * ws_bytes;
      int rc;
      if (L.uniform) {
        rc = b2_filter_lt_u32_dev(ctx, slot, nb, L.batch_len, threshold, d_out, d_end + b0,
                                  d_carry + k + 1, d_carry + k, ws_k, ws_bytes, ctx->s_compute);
      } else {
        rc = b2_filter_lt_u32_ragged_dev(ctx, slot, rel.data(), d_off_k, nb, threshold, d_out,
                                         d_end + b0, d_carry + k + 1, d_carry + k, ws_k, ws_bytes,
                                         ctx->s_compute);
      }
      B2_RETURN_NOT_OK(rc);
      B2_CUDA_OK(ctx, cudaEventRecord(k_end[k], ctx->s_compute));
      work.spans.push_back({kb, k_end[k]});
      B2_CUDA_OK(ctx, cudaMemcpyAsync(const_cast<int64_t*>(h_carry) + k + 1, d_carry + k + 1, 8,
                                      cudaMemcpyDeviceToHost, ctx->s_compute));
      B2_CUDA_OK(ctx, cudaEventRecord(c_ready[k], ctx->s_compute));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_copy_out, k_end[k], 0));
      // keep two chunks queued ahead of the one whose output we wait for
      if (k >= 2) B2_RETURN_NOT_OK(drain(k - 1));
    }
    B2_CUDA_OK(ctx, cudaEventRecord(c_end, ctx->s_copy_in));
    B2_RETURN_NOT_OK(drain(nchunks));
    B2_CUDA_OK(ctx, cudaEventRecord(o_end, ctx->s_copy_out));
    std::vector<int64_t> ends((size_t)nbatches);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(ends.data(), d_end, (size_t)nbatches * 8, cudaMemcpyDeviceToHost,
                                    ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_out));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_in));
    int64_t prev = 0;
    for (int64_t b = 0; b < nbatches; ++b) {
      out_counts[b] = ends[(size_t)b] - prev;
      prev = ends[(size_t)b];
    }
    if (total) *total = (uint64_t)prev;
    float ms = 0;
    cudaEventElapsedTime(&ms, c_begin, c_end);
    tm.copy_to_dev_ms = ms;
    cudaEventElapsedTime(&ms, o_begin, o_end);
    tm.copy_from_dev_ms = ms;
    tm.dev_work_ms = work.total_ms();
    tm.d2h_bytes += nbatches * 8 + (int64_t)nchunks * 8;
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  if (status != B2_OK) return b2_set_error(ctx, status, "b2_filter_lt_u32_host_into", "output buffer too small");
  return B2_OK;
}

// ---- Take ---------------------------------------------------------------------------------
int b2_take_u32_host(b2_ctx* ctx, const uint32_t* const* value_ptrs, const int64_t* value_lens,
                     const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                     uint32_t* const* out_ptrs, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout V, I;
  B2_RETURN_NOT_OK(make_layout(ctx, value_ptrs, value_lens, nbatches, &V));
  B2_RETURN_NOT_OK(make_layout(ctx, idx_ptrs, idx_lens, nbatches, &I));
  B2_REQUIRE(ctx, nbatches == 0 || out_ptrs != nullptr, "out_ptrs is null");
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, idx_lens[b] == 0 || out_ptrs[b] != nullptr, "null output pointer");
    B2_REQUIRE(ctx, idx_lens[b] == 0 || value_lens[b] > 0, "indices into an empty values batch");
  }
  b2_timings tm{};
  if (nbatches > 0 && I.rows() > 0) {
    Scratch sc;
    const bool uniform = V.uniform && I.uniform;
    // chunk on the values layout (the larger side); indices follow the same batch ranges
    const std::vector<int64_t> chunks = make_chunks(V, nbatches);
    gather_begin(ctx, 2 * nbatches, V.rows() + I.rows());
    const size_t nchunks = chunks.size() - 1;
    uint32_t *d_v = nullptr, *d_i = nullptr, *d_o = nullptr;
    int64_t *d_voff = nullptr, *d_ioff = nullptr;
    // grow-only buffers owned by the ctx: no cudaMalloc / cudaFree per call
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 4, (size_t)V.rows() * 4, (void**)&d_v));
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 5, (size_t)I.rows() * 4, (void**)&d_i));
    B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 6, (size_t)I.rows() * 4, (void**)&d_o));
    if (!uniform) {
      char* offs = nullptr;
      B2_RETURN_NOT_OK(b2_ctx_cached(ctx, 7, (size_t)(nbatches + 1) * 16, (void**)&offs));
      d_voff = reinterpret_cast<int64_t*>(offs);
      d_ioff = d_voff + (nbatches + 1);
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_voff, V.off.data(), (size_t)(nbatches + 1) * 8,
                                      cudaMemcpyHostToDevice, ctx->s_compute));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_ioff, I.off.data(), (size_t)(nbatches + 1) * 8,
                                      cudaMemcpyHostToDevice, ctx->s_compute));
    }
    cudaEvent_t c_begin, c_end, o_begin, o_end;
    B2_RETURN_NOT_OK(sc.event(ctx, &c_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &c_end, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &o_begin, true));
    B2_RETURN_NOT_OK(sc.event(ctx, &o_end, true));
    PhaseTimer work;
    B2_CUDA_OK(ctx, cudaEventRecord(c_begin, ctx->s_copy_in));
    for (size_t k = 0; k < nchunks; ++k) {
      cudaEvent_t up_done, kb, ke;
      B2_RETURN_NOT_OK(sc.event(ctx, &up_done, false));
      B2_RETURN_NOT_OK(sc.event(ctx, &kb, true));
      B2_RETURN_NOT_OK(sc.event(ctx, &ke, true));
      const int64_t b0 = chunks[k], b1 = chunks[k + 1];
      B2_RETURN_NOT_OK(upload(ctx, d_v, V, value_ptrs, b0, b1, ctx->s_copy_in, &tm.h2d_bytes));
      B2_RETURN_NOT_OK(upload(ctx, d_i, I, idx_ptrs, b0, b1, ctx->s_copy_in, &tm.h2d_bytes));
      B2_CUDA_OK(ctx, cudaEventRecord(up_done, ctx->s_copy_in));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_compute, up_done, 0));
      B2_CUDA_OK(ctx, cudaEventRecord(kb, ctx->s_compute));
      if (uniform) {
        B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, d_v + V.off[(size_t)b0], V.batch_len,
                                         d_i + I.off[(size_t)b0], I.batch_len, b1 - b0,
                                         d_o + I.off[(size_t)b0], ctx->s_compute));
      } else if (I.off[(size_t)b1] > I.off[(size_t)b0]) {
        // offset tables are global, so run over [0, idx_off[b1]) restricted by pointer shift:
        // the ragged kernel binary-searches the global table, so pass the whole table and range
        B2_RETURN_NOT_OK(b2_take_u32_ragged_dev(ctx, d_v, d_voff, d_i, d_ioff, nbatches,
                                                      I.off[(size_t)b0], I.off[(size_t)b1], d_o,
                                                      ctx->s_compute));
      }
      B2_CUDA_OK(ctx, cudaEventRecord(ke, ctx->s_compute));
      work.spans.push_back({kb, ke});
      // download this chunk's result while the next chunk uploads / computes
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_copy_out, ke, 0));
      if (k == 0) B2_CUDA_OK(ctx, cudaEventRecord(o_begin, ctx->s_copy_out));
      B2_RETURN_NOT_OK(download(ctx, d_o, I, out_ptrs, b0, b1, ctx->s_copy_out, &tm.d2h_bytes));
    }
    B2_CUDA_OK(ctx, cudaEventRecord(c_end, ctx->s_copy_in));
    B2_CUDA_OK(ctx, cudaEventRecord(o_end, ctx->s_copy_out));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_out));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_compute));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_copy_in));
    float ms = 0;
    cudaEventElapsedTime(&ms, c_begin, c_end);
    tm.copy_to_dev_ms = ms;
    cudaEventElapsedTime(&ms, o_begin, o_end);
    tm.copy_from_dev_ms = ms;
    tm.dev_work_ms = work.total_ms();
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}


}  // extern "C"

// 64-bit values: unchunked (upload, one launch per run of equal-shaped batches, download).
extern "C" int b2_take_64_host(b2_ctx* ctx, const void* const* value_ptrs_, const int64_t* value_lens,
                               const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                               void* const* out_ptrs_, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  // the copy helpers move 32-bit words: a batch of n 64-bit values is a batch of 2n words
  const uint32_t* const* value_ptrs = reinterpret_cast<const uint32_t* const*>(value_ptrs_);
  uint32_t* const* out_ptrs = reinterpret_cast<uint32_t* const*>(out_ptrs_);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout V, I;  // in rows
  B2_RETURN_NOT_OK(make_layout(ctx, value_ptrs, value_lens, nbatches, &V));
  B2_RETURN_NOT_OK(make_layout(ctx, idx_ptrs, idx_lens, nbatches, &I));
  B2_REQUIRE(ctx, nbatches == 0 || out_ptrs != nullptr, "out_ptrs is null");
  std::vector<int64_t> vw((size_t)nbatches), ow((size_t)nbatches);
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, idx_lens[b] == 0 || out_ptrs[b] != nullptr, "null output pointer");
    B2_REQUIRE(ctx, idx_lens[b] == 0 || value_lens[b] > 0, "indices into an empty values batch");
    B2_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(value_ptrs[b]) | reinterpret_cast<uintptr_t>(out_ptrs[b])) & 7) == 0,
               "64-bit batches must be 8-byte aligned");
    vw[(size_t)b] = value_lens[b] * 2;
    ow[(size_t)b] = idx_lens[b] * 2;
  }
  Layout VW, OW;  // in 32-bit words
  B2_RETURN_NOT_OK(make_layout(ctx, value_ptrs, vw.data(), nbatches, &VW));
  B2_RETURN_NOT_OK(make_layout(ctx, reinterpret_cast<const uint32_t* const*>(out_ptrs), ow.data(), nbatches, &OW));
  b2_timings tm{};
  if (nbatches > 0 && I.rows() > 0) {
    Scratch sc;
    uint32_t *d_v = nullptr, *d_i = nullptr, *d_o = nullptr;
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_v, (size_t)VW.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_i, (size_t)I.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_o, (size_t)OW.rows() * 4));
    cudaStream_t s = ctx->s_compute;
    gather_begin(ctx, 2 * nbatches, VW.rows() + I.rows());
    B2_RETURN_NOT_OK(upload(ctx, d_v, VW, value_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    B2_RETURN_NOT_OK(upload(ctx, d_i, I, idx_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    // one launch per run of consecutive batches with the same (values, indices) lengths
    for (int64_t b = 0; b < nbatches;) {
      int64_t e = b + 1;
      while (e < nbatches && value_lens[e] == value_lens[b] && idx_lens[e] == idx_lens[b]) ++e;
      if (idx_lens[b] > 0) {
        B2_RETURN_NOT_OK(b2_take_64_dev(ctx, d_v + VW.off[(size_t)b], value_lens[b], d_i + I.off[(size_t)b],
                                        idx_lens[b], e - b, d_o + OW.off[(size_t)b], s));
      }
      b = e;
    }
    B2_RETURN_NOT_OK(download(ctx, d_o, OW, out_ptrs, 0, nbatches, s, &tm.d2h_bytes));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

// ---- nullable columns ---------------------------------------------------------
// Straightforward (unchunked) host entry points: the per-batch Arrow validity bitmaps are packed
// into ONE bitmap over the packed device column on the host, uploaded with the values, and the
// nullable kernels run on the whole column.
namespace {

void set_bits(uint8_t* dst, int64_t d0, int64_t n) {
  for (int64_t i = 0; i < n;) {
    const int64_t d = d0 + i;
    if ((d & 7) == 0 && n - i >= 8) {
      const int64_t nbytes = (n - i) >> 3;
      memset(dst + (d >> 3), 0xff, (size_t)nbytes);
      i += nbytes << 3;
    } else {
      dst[d >> 3] |= (uint8_t)(1u << (d & 7));
      ++i;
    }
  }
}

// dst bits [d0, d0+n) = src bits [s0, s0+n); dst bits are zero beforehand.
void copy_bits(uint8_t* dst, int64_t d0, const uint8_t* src, int64_t s0, int64_t n) {
  int64_t i = 0;
  if (((d0 | s0) & 7) == 0) {
    const int64_t nbytes = n >> 3;
    memcpy(dst + (d0 >> 3), src + (s0 >> 3), (size_t)nbytes);
    i = nbytes << 3;
  }
  for (; i < n; ++i) {
    const int64_t s = s0 + i, d = d0 + i;
    if ((src[s >> 3] >> (s & 7)) & 1) dst[d >> 3] |= (uint8_t)(1u << (d & 7));
  }
}

// One bitmap over the packed column (bit L.off[b] + r = row r of batch b), padded to 4 bytes.
// valid_ptrs == NULL or valid_ptrs[b] == NULL: batch b has no nulls. Returns true if any batch
// brought a bitmap.
bool pack_validity(std::vector<uint8_t>* dst, const Layout& L, const uint8_t* const* valid_ptrs,
                   const int64_t* bit_offsets, int64_t nbatches) {
  dst->assign(b2_align_up((size_t)((L.rows() + 7) >> 3), 4) + 4, 0);
  bool any = false;
  for (int64_t b = 0; b < nbatches; ++b) {
    const int64_t n = L.off[(size_t)b + 1] - L.off[(size_t)b];
    if (valid_ptrs && valid_ptrs[b]) {
      any = true;
      copy_bits(dst->data(), L.off[(size_t)b], valid_ptrs[b], bit_offsets ? bit_offsets[b] : 0, n);
    } else {
      set_bits(dst->data(), L.off[(size_t)b], n);
    }
  }
  return any;
}

}  // namespace

extern "C" {

int b2_aggr_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                     const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                     b2_aggr_u32* out, b2_timings* timings) {
  return b2_aggr_32_host(ctx, reinterpret_cast<const void* const*>(batch_ptrs), valid_ptrs, valid_bit_offsets,
                         batch_lens, nbatches, B2_U32, out, timings);
}

int b2_aggr_32_host(b2_ctx* ctx, const void* const* batch_ptrs_, const uint8_t* const* valid_ptrs,
                    const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches, int dtype,
                    b2_aggr_u32* out, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const uint32_t* const* batch_ptrs = reinterpret_cast<const uint32_t* const*>(batch_ptrs_);
  B2_REQUIRE(ctx, out != nullptr, "out is null");
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout L;
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  b2_timings tm{};
  Scratch sc;
  std::vector<uint8_t> bits;
  const bool nullable = pack_validity(&bits, L, valid_ptrs, valid_bit_offsets, nbatches);
  uint32_t* d_col = nullptr;
  uint8_t* d_valid = nullptr;
  b2_aggr_u32* d_out = nullptr;
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_col, (size_t)L.rows() * 4));
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_out, sizeof(b2_aggr_u32)));
  cudaStream_t s = ctx->s_compute;
  gather_begin(ctx, nbatches, L.rows());
  B2_RETURN_NOT_OK(upload(ctx, d_col, L, batch_ptrs, 0, nbatches, s, &tm.h2d_bytes));
  if (nullable) {
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_valid, bits.size()));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(d_valid, bits.data(), bits.size(), cudaMemcpyHostToDevice, s));
    tm.h2d_bytes += (int64_t)bits.size();
  }
  B2_RETURN_NOT_OK(b2_aggr_32_dev(ctx, d_col, dtype, d_valid, L.rows(), d_out, s));
  B2_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, sizeof(b2_aggr_u32), cudaMemcpyDeviceToHost, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  tm.d2h_bytes = sizeof(b2_aggr_u32);
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_aggr_64_host(b2_ctx* ctx, const void* const* batch_ptrs_, const uint8_t* const* valid_ptrs,
                    const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches, int dtype,
                    b2_aggr_u64* out, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  // the upload machinery moves 32-bit words: a batch of n 64-bit values is a batch of 2n words
  const uint32_t* const* batch_ptrs = reinterpret_cast<const uint32_t* const*>(batch_ptrs_);
  B2_REQUIRE(ctx, out != nullptr, "out is null");
  B2_REQUIRE(ctx, nbatches >= 0 && (nbatches == 0 || batch_lens), "bad batch table");
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout L, W;  // rows (validity) and words (upload)
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  std::vector<int64_t> wlens((size_t)nbatches);
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(batch_ptrs[b]) & 7) == 0, "batches must be 8-byte aligned");
    wlens[(size_t)b] = batch_lens[b] * 2;
  }
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, wlens.data(), nbatches, &W));
  b2_timings tm{};
  Scratch sc;
  std::vector<uint8_t> bits;
  const bool nullable = pack_validity(&bits, L, valid_ptrs, valid_bit_offsets, nbatches);
  uint32_t* d_col = nullptr;
  uint8_t* d_valid = nullptr;
  b2_aggr_u64* d_out = nullptr;
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_col, (size_t)W.rows() * 4));
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_out, sizeof(b2_aggr_u64)));
  cudaStream_t s = ctx->s_compute;
  gather_begin(ctx, nbatches, W.rows());
  B2_RETURN_NOT_OK(upload(ctx, d_col, W, batch_ptrs, 0, nbatches, s, &tm.h2d_bytes));
  if (nullable) {
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_valid, bits.size()));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(d_valid, bits.data(), bits.size(), cudaMemcpyHostToDevice, s));
    tm.h2d_bytes += (int64_t)bits.size();
  }
  B2_RETURN_NOT_OK(b2_aggr_64_dev(ctx, d_col, dtype, d_valid, L.rows(), d_out, s));
  B2_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, sizeof(b2_aggr_u64), cudaMemcpyDeviceToHost, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  tm.d2h_bytes = sizeof(b2_aggr_u64);
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

static int filter_typed_host_into(b2_ctx* ctx, int dtype, const uint32_t* const* batch_ptrs,
                                  const uint8_t* const* valid_ptrs, const int64_t* valid_bit_offsets,
                                  const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                                  uint32_t* out, int64_t out_capacity, int64_t* out_counts,
                                  uint64_t* total, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, total != nullptr && out_capacity >= 0, "bad result arguments");
  B2_REQUIRE(ctx, nbatches == 0 || out_counts != nullptr, "out_counts is null");
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout L;
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  *total = 0;
  b2_timings tm{};
  if (L.rows() > 0) {
    Scratch sc;
    std::vector<uint8_t> bits;
    const bool nullable = pack_validity(&bits, L, valid_ptrs, valid_bit_offsets, nbatches);
    uint32_t *d_col = nullptr, *d_out = nullptr;
    uint8_t* d_valid = nullptr;
    int64_t* d_end = nullptr;
    void* d_ws = nullptr;
    const size_t ws_bytes = L.uniform ? b2_filter_ws_bytes(nbatches, L.batch_len)
                                      : b2_filter_ragged_ws_bytes(L.off.data(), nbatches);
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_col, (size_t)L.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_out, (size_t)L.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_end, (size_t)(nbatches + 1) * 8));
    B2_RETURN_NOT_OK(sc.alloc(ctx, &d_ws, ws_bytes));
    cudaStream_t s = ctx->s_compute;
    int64_t* d_off = nullptr;
    if (!L.uniform) {  // ragged batches: the kernel takes its tile geometry from the offset table
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_off, (size_t)(nbatches + 1) * 8));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_off, L.off.data(), (size_t)(nbatches + 1) * 8, cudaMemcpyHostToDevice, s));
    }
    gather_begin(ctx, nbatches, L.rows());
    B2_RETURN_NOT_OK(upload(ctx, d_col, L, batch_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    if (nullable) {
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_valid, bits.size()));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_valid, bits.data(), bits.size(), cudaMemcpyHostToDevice, s));
      tm.h2d_bytes += (int64_t)bits.size();
    }
    if (L.uniform)
      B2_RETURN_NOT_OK(b2_filter_lt_32_dev(ctx, d_col, dtype, threshold, d_valid, nbatches, L.batch_len, d_out,
                                           d_end, d_end + nbatches, nullptr, d_ws, ws_bytes, s));
    else
      B2_RETURN_NOT_OK(b2_filter_lt_32_ragged_dev(ctx, d_col, dtype, threshold, d_valid, L.off.data(), d_off,
                                                  nbatches, d_out, d_end, d_end + nbatches, nullptr, d_ws,
                                                  ws_bytes, s));
    std::vector<int64_t> end((size_t)nbatches + 1);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(end.data(), d_end, (size_t)(nbatches + 1) * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    for (int64_t b = 0; b < nbatches; ++b) out_counts[b] = end[(size_t)b] - (b ? end[(size_t)b - 1] : 0);
    *total = (uint64_t)end[(size_t)nbatches];
    tm.d2h_bytes = (nbatches + 1) * 8;
    if ((int64_t)*total > out_capacity)
      return b2_set_error(ctx, B2_ERR_OVERFLOW, "nullable filter", "result exceeds out_capacity");
    if (*total > 0) {
      B2_REQUIRE(ctx, out != nullptr, "out is null");
      B2_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, (size_t)*total * 4, cudaMemcpyDeviceToHost, s));
      B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
      tm.d2h_bytes += (int64_t)*total * 4;
    }
  } else {
    for (int64_t b = 0; b < nbatches; ++b) out_counts[b] = 0;
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_filter_lt_u32_nullable_host_into(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                                        const uint8_t* const* valid_ptrs, const int64_t* valid_bit_offsets,
                                        const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                                        uint32_t* out, int64_t out_capacity, int64_t* out_counts,
                                        uint64_t* total, b2_timings* timings) {
  return filter_typed_host_into(ctx, B2_U32, batch_ptrs, valid_ptrs, valid_bit_offsets, batch_lens, nbatches,
                                threshold, out, out_capacity, out_counts, total, timings);
}

int b2_filter_lt_32_host_into(b2_ctx* ctx, const void* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                              const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                              int dtype, uint32_t threshold_bits, void* out, int64_t out_capacity,
                              int64_t* out_counts, uint64_t* total, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, dtype == B2_U32 || dtype == B2_I32 || dtype == B2_F32, "dtype must be B2_U32, B2_I32 or B2_F32");
  return filter_typed_host_into(ctx, dtype, reinterpret_cast<const uint32_t* const*>(batch_ptrs), valid_ptrs,
                                valid_bit_offsets, batch_lens, nbatches, threshold_bits,
                                static_cast<uint32_t*>(out), out_capacity, out_counts, total, timings);
}

int b2_filter_lt_64_host_into(b2_ctx* ctx, const void* const* batch_ptrs_, const uint8_t* const* valid_ptrs,
                              const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                              int dtype, uint64_t threshold_bits, void* out, int64_t out_capacity,
                              int64_t* out_counts, uint64_t* total, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, dtype == B2_U64 || dtype == B2_I64 || dtype == B2_F64, "dtype must be B2_U64, B2_I64 or B2_F64");
  B2_REQUIRE(ctx, total != nullptr && out_capacity >= 0, "bad result arguments");
  B2_REQUIRE(ctx, nbatches >= 0 && (nbatches == 0 || (batch_lens && out_counts)), "bad batch table");
  // the upload machinery moves 32-bit words: a batch of n 64-bit values is a batch of 2n words
  const uint32_t* const* batch_ptrs = reinterpret_cast<const uint32_t* const*>(batch_ptrs_);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout L, W;  // rows (validity, boundaries) and words (upload)
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, batch_lens, nbatches, &L));
  std::vector<int64_t> wlens((size_t)nbatches);
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(batch_ptrs[b]) & 7) == 0, "batches must be 8-byte aligned");
    wlens[(size_t)b] = batch_lens[b] * 2;
  }
  B2_RETURN_NOT_OK(make_layout(ctx, batch_ptrs, wlens.data(), nbatches, &W));
  *total = 0;
  b2_timings tm{};
  if (L.rows() > 0) {
    Scratch sc;
    std::vector<uint8_t> bits;
    const bool nullable = pack_validity(&bits, L, valid_ptrs, valid_bit_offsets, nbatches);
    uint32_t* d_col = nullptr;
    uint64_t* d_out = nullptr;
    uint8_t* d_valid = nullptr;
    int64_t *d_end = nullptr, *d_off = nullptr;
    void* d_ws = nullptr;
    const size_t ws_bytes = b2_filter_64_ws_bytes(L.rows());
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_col, (size_t)W.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_out, (size_t)L.rows() * 8));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_end, (size_t)(nbatches + 1) * 8));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_off, (size_t)(nbatches + 1) * 8));
    B2_RETURN_NOT_OK(sc.alloc(ctx, &d_ws, ws_bytes));
    cudaStream_t s = ctx->s_compute;
    B2_CUDA_OK(ctx, cudaMemcpyAsync(d_off, L.off.data(), (size_t)(nbatches + 1) * 8, cudaMemcpyHostToDevice, s));
    gather_begin(ctx, nbatches, W.rows());
    B2_RETURN_NOT_OK(upload(ctx, d_col, W, batch_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    if (nullable) {
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_valid, bits.size()));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_valid, bits.data(), bits.size(), cudaMemcpyHostToDevice, s));
      tm.h2d_bytes += (int64_t)bits.size();
    }
    B2_RETURN_NOT_OK(b2_filter_lt_64_dev(ctx, d_col, dtype, threshold_bits, d_valid, L.rows(), d_off, nbatches, 0,
                                         d_out, d_end, d_end + nbatches, d_ws, ws_bytes, s));
    std::vector<int64_t> end((size_t)nbatches + 1);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(end.data(), d_end, (size_t)(nbatches + 1) * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    for (int64_t b = 0; b < nbatches; ++b) out_counts[b] = end[(size_t)b] - (b ? end[(size_t)b - 1] : 0);
    *total = (uint64_t)end[(size_t)nbatches];
    tm.d2h_bytes = (nbatches + 1) * 8;
    if ((int64_t)*total > out_capacity)
      return b2_set_error(ctx, B2_ERR_OVERFLOW, "64-bit filter", "result exceeds out_capacity");
    if (*total > 0) {
      B2_REQUIRE(ctx, out != nullptr, "out is null");
      B2_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, (size_t)*total * 8, cudaMemcpyDeviceToHost, s));
      B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
      tm.d2h_bytes += (int64_t)*total * 8;
    }
  } else {
    for (int64_t b = 0; b < nbatches; ++b) out_counts[b] = 0;
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

// ---- join with NULLABLE keys --------------------------------------------------
// Arrow's inner hash join (the reference's oracle, join_native.cc:31-36) never matches a null key,
// on either side; the DPU path has no bitmaps at all. Rows with a null key are dropped ON THE DEVICE
// before the join, with kernels that already exist: iota -> nullable filter (keeps the row numbers
// whose key is valid) -> take(key), take(payload) -> the plain join. A side without nulls skips it.
namespace {
int drop_null_keys(b2_ctx* ctx, Scratch* sc, const uint32_t* d_key, const uint32_t* d_pay, int64_t n,
                   const std::vector<uint8_t>& bits, const uint32_t** key_out, const uint32_t** pay_out,
                   int64_t* n_out, cudaStream_t s) {
  B2_REQUIRE(ctx, n < (1ll << 32) - 1, "row numbers travel as uint32");
  uint8_t* d_valid = nullptr;
  uint32_t *d_iota = nullptr, *d_idx = nullptr, *d_k = nullptr, *d_p = nullptr;
  int64_t *d_end = nullptr, *d_total = nullptr;
  void* d_ws = nullptr;
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_valid, bits.size()));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_iota, (size_t)n * 4));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_idx, (size_t)n * 4));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_end, 8));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_total, 8));
  const size_t ws_bytes = b2_filter_ws_bytes(1, n);
  B2_RETURN_NOT_OK(sc->alloc(ctx, &d_ws, ws_bytes));
  B2_CUDA_OK(ctx, cudaMemcpyAsync(d_valid, bits.data(), bits.size(), cudaMemcpyHostToDevice, s));
  B2_RETURN_NOT_OK(b2_iota_u32_dev(ctx, 0, n, d_iota, s));
  // every row number is below 0xffffffff, so the predicate keeps exactly the rows whose key is valid
  B2_RETURN_NOT_OK(b2_filter_lt_u32_nullable_dev(ctx, d_iota, d_valid, 1, n, 0xffffffffu, d_idx, d_end, d_total,
                                                 nullptr, d_ws, ws_bytes, s));
  int64_t kept = 0;
  B2_CUDA_OK(ctx, cudaMemcpyAsync(&kept, d_total, 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_k, (size_t)kept * 4));
  B2_RETURN_NOT_OK(sc->alloc(ctx, (void**)&d_p, (size_t)kept * 4));
  if (kept > 0) {
    B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, d_key, n, d_idx, kept, 1, d_k, s));
    B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, d_pay, n, d_idx, kept, 1, d_p, s));
  }
  *key_out = d_k;
  *pay_out = d_p;
  *n_out = kept;
  return B2_OK;
}
}  // namespace

int b2_join_u32_nullable_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const uint8_t* const* l_key_valid_ptrs,
                              const int64_t* l_key_valid_bit_offsets, const int64_t* l_lens, int64_t nl_batches,
                              const uint32_t* const* r_ptrs, const uint8_t* const* r_key_valid_ptrs,
                              const int64_t* r_key_valid_bit_offsets, const int64_t* r_lens, int64_t nr_batches,
                              uint64_t* out_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, out_rows != nullptr, "out_rows is null");
  Layout LL, LR;
  B2_RETURN_NOT_OK(make_layout(ctx, l_ptrs, l_lens, nl_batches, &LL));
  B2_RETURN_NOT_OK(make_layout(ctx, r_ptrs, r_lens, nr_batches, &LR));
  std::vector<uint8_t> lbits, rbits;
  const bool lnull = pack_validity(&lbits, LL, l_key_valid_ptrs, l_key_valid_bit_offsets, nl_batches);
  const bool rnull = pack_validity(&rbits, LR, r_key_valid_ptrs, r_key_valid_bit_offsets, nr_batches);
  if (!lnull && !rnull)  // no bitmap anywhere: the plain join
    return b2_join_u32_host(ctx, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches, out_rows, timings);
  cudaStream_t s = ctx->s_compute;
  Scratch sc;
  b2_timings tm{};
  const int64_t nl = LL.rows(), nr = LR.rows();
  uint32_t *d_fk, *d_y, *d_pk, *d_x;
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_fk, (size_t)nl * 4));
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_y, (size_t)nl * 4));
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_pk, (size_t)nr * 4));
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_x, (size_t)nr * 4));
  cudaEvent_t e0, e1, e2;
  B2_RETURN_NOT_OK(sc.event(ctx, &e0, true));
  B2_RETURN_NOT_OK(sc.event(ctx, &e1, true));
  B2_RETURN_NOT_OK(sc.event(ctx, &e2, true));
  B2_CUDA_OK(ctx, cudaEventRecord(e0, s));
  gather_begin(ctx, 2 * (nl_batches + nr_batches), 2 * (nl + nr));
  B2_RETURN_NOT_OK(upload(ctx, d_fk, LL, l_ptrs, 0, nl_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload(ctx, d_y, LL, l_ptrs + nl_batches, 0, nl_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload(ctx, d_pk, LR, r_ptrs, 0, nr_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload(ctx, d_x, LR, r_ptrs + nr_batches, 0, nr_batches, s, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(e1, s));
  const uint32_t *k_l = d_fk, *p_l = d_y, *k_r = d_pk, *p_r = d_x;
  int64_t ml = nl, mr = nr;
  if (lnull) B2_RETURN_NOT_OK(drop_null_keys(ctx, &sc, d_fk, d_y, nl, lbits, &k_l, &p_l, &ml, s));
  if (rnull) B2_RETURN_NOT_OK(drop_null_keys(ctx, &sc, d_pk, d_x, nr, rbits, &k_r, &p_r, &mr, s));
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(ml, mr);
  B2_RETURN_NOT_OK(sc.alloc(ctx, &d_ws, ws_bytes));
  uint64_t* d_rows = nullptr;
  B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_rows, 8));
  int64_t cap = ml;
  uint32_t *o_fk = nullptr, *o_y = nullptr, *o_x = nullptr;
  uint64_t rows = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2_RETURN_NOT_OK(dev_alloc(ctx, (void**)&o_fk, (size_t)cap * 4));
    B2_RETURN_NOT_OK(dev_alloc(ctx, (void**)&o_y, (size_t)cap * 4));
    B2_RETURN_NOT_OK(dev_alloc(ctx, (void**)&o_x, (size_t)cap * 4));
    int rc = b2_join_u32_dev(ctx, k_l, p_l, ml, k_r, p_r, mr, o_fk, o_y, o_x, cap, d_rows, 0, d_ws, ws_bytes, s);
    if (rc == B2_OK && (cudaMemcpyAsync(&rows, d_rows, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                        cudaStreamSynchronize(s) != cudaSuccess))
      rc = b2_set_error(ctx, B2_ERR_CUDA, "join row count", cudaGetErrorString(cudaGetLastError()));
    if (rc == B2_OK && rows == ~0ull) rc = b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
    if (rc == B2_OK && (int64_t)rows > cap && attempt == 1)
      rc = b2_set_error(ctx, B2_ERR_OVERFLOW, "join", "output larger than reported");
    if (rc != B2_OK || (int64_t)rows > cap) {
      for (uint32_t* p : {o_fk, o_y, o_x}) b2_dev_free(ctx, p);
      if (rc != B2_OK) return rc;
      cap = (int64_t)rows;
      continue;
    }
    break;
  }
  B2_CUDA_OK(ctx, cudaEventRecord(e2, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kJoin;
  pend->d_fk = o_fk;
  pend->d_y = o_y;
  pend->d_x = o_x;
  pend->rows = rows;
  for (uint32_t* p : {o_fk, o_y, o_x}) pend->dev.push_back(p);
  ctx->pending = pend;
  *out_rows = rows;
  float up = 0, work = 0;
  cudaEventElapsedTime(&up, e0, e1);
  cudaEventElapsedTime(&work, e1, e2);
  tm.copy_to_dev_ms = up;
  tm.dev_work_ms = work;
  tm.h2d_bytes += (int64_t)(lnull ? lbits.size() : 0) + (int64_t)(rnull ? rbits.size() : 0);
  tm.d2h_bytes = 8;
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_take_u32_nullable_host(b2_ctx* ctx, const uint32_t* const* value_ptrs,
                              const uint8_t* const* value_valid_ptrs, const int64_t* value_valid_bit_offsets,
                              const int64_t* value_lens, const uint32_t* const* idx_ptrs,
                              const uint8_t* const* idx_valid_ptrs, const int64_t* idx_valid_bit_offsets,
                              const int64_t* idx_lens, int64_t nbatches, uint32_t* const* out_ptrs,
                              uint8_t* const* out_valid_ptrs, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  Layout LV, LI;
  B2_RETURN_NOT_OK(make_layout(ctx, value_ptrs, value_lens, nbatches, &LV));
  B2_RETURN_NOT_OK(make_layout(ctx, idx_ptrs, idx_lens, nbatches, &LI));
  if (!LV.uniform || !LI.uniform)
    return b2_set_error(ctx, B2_ERR_UNSUPPORTED, "nullable take", "batches must have equal lengths");
  B2_REQUIRE(ctx, nbatches == 0 || (out_ptrs && out_valid_ptrs), "null output tables");
  b2_timings tm{};
  if (LI.rows() > 0) {
    Scratch sc;
    std::vector<uint8_t> vbits, ibits;
    const bool vnull = pack_validity(&vbits, LV, value_valid_ptrs, value_valid_bit_offsets, nbatches);
    const bool inull = pack_validity(&ibits, LI, idx_valid_ptrs, idx_valid_bit_offsets, nbatches);
    uint32_t *d_val = nullptr, *d_idx = nullptr, *d_out = nullptr;
    uint8_t *d_vv = nullptr, *d_iv = nullptr, *d_ov = nullptr;
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_val, (size_t)LV.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_idx, (size_t)LI.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_out, (size_t)LI.rows() * 4));
    B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_ov, ibits.size()));
    cudaStream_t s = ctx->s_compute;
    gather_begin(ctx, 2 * nbatches, LV.rows() + LI.rows());
    B2_RETURN_NOT_OK(upload(ctx, d_val, LV, value_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    B2_RETURN_NOT_OK(upload(ctx, d_idx, LI, idx_ptrs, 0, nbatches, s, &tm.h2d_bytes));
    if (vnull) {
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_vv, vbits.size()));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_vv, vbits.data(), vbits.size(), cudaMemcpyHostToDevice, s));
      tm.h2d_bytes += (int64_t)vbits.size();
    }
    if (inull) {
      B2_RETURN_NOT_OK(sc.alloc(ctx, (void**)&d_iv, ibits.size()));
      B2_CUDA_OK(ctx, cudaMemcpyAsync(d_iv, ibits.data(), ibits.size(), cudaMemcpyHostToDevice, s));
      tm.h2d_bytes += (int64_t)ibits.size();
    }
    B2_RETURN_NOT_OK(b2_take_u32_nullable_dev(ctx, d_val, d_vv, LV.batch_len, d_idx, d_iv, LI.batch_len,
                                              nbatches, d_out, d_ov, s));
    B2_RETURN_NOT_OK(download(ctx, d_out, LI, out_ptrs, 0, nbatches, s, &tm.d2h_bytes));
    std::vector<uint8_t> obits(ibits.size());
    B2_CUDA_OK(ctx, cudaMemcpyAsync(obits.data(), d_ov, obits.size(), cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    tm.d2h_bytes += (int64_t)obits.size();
    for (int64_t b = 0; b < nbatches; ++b) {  // per-batch result bitmaps, bit offset 0
      const int64_t n = LI.off[(size_t)b + 1] - LI.off[(size_t)b];
      memset(out_valid_ptrs[b], 0, (size_t)((n + 7) >> 3));
      copy_bits(out_valid_ptrs[b], 0, obits.data(), LI.off[(size_t)b], n);
    }
  }
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

}  // extern "C"
