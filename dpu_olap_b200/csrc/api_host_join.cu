// api_host_join.cu — host-buffer entry points of Partition and Join: what PartitionDpu::Run
// (host/partition/partition_dpu.cc:31-135) and JoinDpu::Run (host/join/join_dpu.cc:158-400) did
// with DpuSet transfers. Columns are uploaded once, the whole operator runs on the device, and
// the result stays there until the caller — who can only size its buffers afterwards — fetches it.
#include <algorithm>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "pending.h"

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

int ensure_streams(b2_ctx* ctx) {
  B2_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->s_compute) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  if (!ctx->s_copy_in) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_in, cudaStreamNonBlocking));
  if (!ctx->s_copy_out) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_copy_out, cudaStreamNonBlocking));
  return B2_OK;
}

struct DevBufs {  // returns everything not released to the ctx's recycling pool
  b2_ctx* owner = nullptr;
  std::vector<void*> bufs;
  ~DevBufs() {
    if (owner) {  // error paths leave with work in flight: nothing it uses is recycled before it has drained
      b2_device_scope sc(owner);
      for (cudaStream_t st : {owner->s_copy_in, owner->s_compute, owner->s_copy_out})
        if (st) cudaStreamSynchronize(st);
    }
    for (void* p : bufs)
      if (p) b2_dev_free(owner, p);
  }
  int alloc(b2_ctx* ctx, void** p, size_t bytes) {
    owner = ctx;
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, p, bytes ? bytes : 256));
    bufs.push_back(*p);
    return B2_OK;
  }
  void release(void* p) {
    for (auto& q : bufs)
      if (q == p) q = nullptr;
  }
};

// Upload one column given as nbatches host buffers into a packed device column.
int upload_column(b2_ctx* ctx, uint32_t* d_col, const uint32_t* const* ptrs, const int64_t* lens,
                  int64_t nbatches, cudaStream_t s, int64_t* bytes) {
  int64_t off = 0, b = 0;
  while (b < nbatches) {
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

int total_rows(b2_ctx* ctx, const int64_t* lens, int64_t nbatches, int64_t* n) {
  *n = 0;
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, lens[b] >= 0, "negative batch length");
    *n += lens[b];
  }
  return B2_OK;
}

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  int init(b2_ctx* ctx) {
    B2_CUDA_OK(ctx, cudaEventCreate(&a));
    B2_CUDA_OK(ctx, cudaEventCreate(&b));
    return B2_OK;
  }
  double ms() const {
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    return t;
  }
};

}  // namespace

extern "C" {

int b2_partition_u32_host(b2_ctx* ctx, const uint32_t* const* col_batch_ptrs,
                          const int64_t* batch_lens, int64_t nbatches, int ncols, int key_col,
                          int nparts, int64_t* part_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nbatches >= 0 && ncols >= 1 && ncols <= 16, "bad shape");
  B2_REQUIRE(ctx, key_col >= 0 && key_col < ncols, "key column out of range");
  B2_REQUIRE(ctx, nparts >= 1 && (nparts & (nparts - 1)) == 0, "nparts must be a power of two");
  B2_REQUIRE(ctx, part_rows != nullptr, "part_rows is null");
  B2_REQUIRE(ctx, nbatches == 0 || (col_batch_ptrs && batch_lens), "null batch table");
  int64_t n = 0;
  B2_RETURN_NOT_OK(total_rows(ctx, batch_lens, nbatches, &n));
  b2_timings tm{};
  DevBufs bufs;
  EventPair up, work;
  B2_RETURN_NOT_OK(up.init(ctx));
  B2_RETURN_NOT_OK(work.init(ctx));
  cudaStream_t s = ctx->s_compute;
  std::vector<uint32_t*> d_in((size_t)ncols), d_out((size_t)ncols);
  B2_CUDA_OK(ctx, cudaEventRecord(up.a, s));
  for (int c = 0; c < ncols; ++c) {
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_in[(size_t)c], (size_t)n * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_out[(size_t)c], (size_t)n * 4));
    B2_RETURN_NOT_OK(upload_column(ctx, d_in[(size_t)c], col_batch_ptrs + (size_t)c * nbatches,
                                   batch_lens, nbatches, s, &tm.h2d_bytes));
  }
  B2_CUDA_OK(ctx, cudaEventRecord(up.b, s));
  // the C ABI wants the key column first
  std::vector<const uint32_t*> in_order;
  std::vector<uint32_t*> out_order;
  in_order.push_back(d_in[(size_t)key_col]);
  out_order.push_back(d_out[(size_t)key_col]);
  for (int c = 0; c < ncols; ++c)
    if (c != key_col) {
      in_order.push_back(d_in[(size_t)c]);
      out_order.push_back(d_out[(size_t)c]);
    }
  void* d_ws = nullptr;
  int64_t* d_off = nullptr;
  const size_t ws_bytes = b2_partition_ws_bytes(n, nparts);
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_ws, ws_bytes));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_off, ((size_t)nparts + 1) * 8));
  B2_CUDA_OK(ctx, cudaEventRecord(work.a, s));
  B2_RETURN_NOT_OK(b2_partition_u32_dev(ctx, in_order.data(), out_order.data(), ncols, n, nparts, 0,
                                        d_off, d_ws, ws_bytes, s));
  B2_CUDA_OK(ctx, cudaEventRecord(work.b, s));
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kPartition;
  pend->ncols = ncols;
  pend->part_off.assign((size_t)nparts + 1, 0);
  ctx->pending = pend;
  B2_CUDA_OK(ctx, cudaMemcpyAsync(pend->part_off.data(), d_off, ((size_t)nparts + 1) * 8,
                                  cudaMemcpyDeviceToHost, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  for (int p = 0; p < nparts; ++p) part_rows[p] = pend->part_off[(size_t)p + 1] - pend->part_off[(size_t)p];
  for (int c = 0; c < ncols; ++c) {
    pend->d_cols.push_back(d_out[(size_t)c]);
    pend->dev.push_back(d_out[(size_t)c]);
    bufs.release(d_out[(size_t)c]);
  }
  tm.copy_to_dev_ms = up.ms();
  tm.dev_work_ms = work.ms();
  tm.d2h_bytes = ((int64_t)nparts + 1) * 8;
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_partition_fetch_host(b2_ctx* ctx, uint32_t* const* out_ptrs, int nparts, int ncols,
                            b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  b2_pending* pend = ctx->pending;
  if (!pend || pend->kind != b2_pending::kPartition)
    return b2_set_error(ctx, B2_ERR_INVALID, "b2_partition_fetch_host", "no pending partition result");
  B2_REQUIRE(ctx, nparts + 1 == (int)pend->part_off.size() && ncols == pend->ncols,
             "shape differs from the run");
  B2_REQUIRE(ctx, out_ptrs != nullptr, "out_ptrs is null");
  b2_timings tm{};
  EventPair ev;
  B2_RETURN_NOT_OK(ev.init(ctx));
  cudaStream_t s = ctx->s_copy_out;
  B2_CUDA_OK(ctx, cudaEventRecord(ev.a, s));
  for (int p = 0; p < nparts; ++p) {
    const int64_t r0 = pend->part_off[(size_t)p], rows = pend->part_off[(size_t)p + 1] - r0;
    if (rows == 0) continue;
    for (int c = 0; c < ncols; ++c) {
      uint32_t* dst = out_ptrs[(size_t)p * ncols + c];
      B2_REQUIRE(ctx, dst != nullptr, "null output pointer for a non-empty partition");
      B2_CUDA_OK(ctx, cudaMemcpyAsync(dst, pend->d_cols[(size_t)c] + r0, (size_t)rows * 4,
                                      cudaMemcpyDeviceToHost, s));
      tm.d2h_bytes += rows * 4;
    }
  }
  B2_CUDA_OK(ctx, cudaEventRecord(ev.b, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  tm.copy_from_dev_ms = ev.ms();
  tm.total_ms = ms_since(t0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_join_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens,
                     int64_t nl_batches, const uint32_t* const* r_ptrs, const int64_t* r_lens,
                     int64_t nr_batches, uint64_t* out_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nl_batches >= 0 && nr_batches >= 0, "negative batch count");
  B2_REQUIRE(ctx, out_rows != nullptr, "out_rows is null");
  B2_REQUIRE(ctx, nl_batches == 0 || (l_ptrs && l_lens), "null left batch table");
  B2_REQUIRE(ctx, nr_batches == 0 || (r_ptrs && r_lens), "null right batch table");
  int64_t nl = 0, nr = 0;
  B2_RETURN_NOT_OK(total_rows(ctx, l_lens, nl_batches, &nl));
  B2_RETURN_NOT_OK(total_rows(ctx, r_lens, nr_batches, &nr));
  b2_timings tm{};
  DevBufs bufs;
  EventPair up, work;
  B2_RETURN_NOT_OK(up.init(ctx));
  B2_RETURN_NOT_OK(work.init(ctx));
  cudaStream_t s = ctx->s_compute;
  uint32_t *d_fk, *d_y, *d_pk, *d_x;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_fk, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_y, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_pk, (size_t)nr * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_x, (size_t)nr * 4));
  // The build side goes up first, on the copy stream; its radix passes run on the compute stream
  // while the probe side is still crossing PCIe (the reference moves each column four times and
  // strictly one after the other, join_dpu.cc:168-400).
  cudaStream_t sc = ctx->s_copy_in;
  EventPair landed;  // a: build side on the device, b: probe side on the device
  B2_RETURN_NOT_OK(landed.init(ctx));
  B2_CUDA_OK(ctx, cudaEventRecord(up.a, sc));
  B2_RETURN_NOT_OK(upload_column(ctx, d_pk, r_ptrs, r_lens, nr_batches, sc, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_x, r_ptrs + nr_batches, r_lens, nr_batches, sc, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(landed.a, sc));
  B2_RETURN_NOT_OK(upload_column(ctx, d_fk, l_ptrs, l_lens, nl_batches, sc, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_y, l_ptrs + nl_batches, l_lens, nl_batches, sc, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(landed.b, sc));
  B2_CUDA_OK(ctx, cudaEventRecord(up.b, sc));

  // PK-FK joins produce at most nl rows; duplicate build keys can produce more, in which case the
  // join is re-run with the exact capacity it reported.
  int64_t cap = nl;
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(nl, nr);
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_ws, ws_bytes));
  uint64_t* d_rows = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rows, 8));
  uint32_t *o_fk = nullptr, *o_y = nullptr, *o_x = nullptr;
  uint64_t rows = 0;
  B2_CUDA_OK(ctx, cudaStreamWaitEvent(s, landed.a, 0));
  B2_CUDA_OK(ctx, cudaEventRecord(work.a, s));
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_fk, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_y, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_x, (size_t)cap * 4));
    if (attempt == 0) {
      B2_RETURN_NOT_OK(b2_join_u32_phased_dev(ctx, d_fk, d_y, nl, d_pk, d_x, nr, o_fk, o_y, o_x, cap, d_rows, 0, 1,
                                              d_ws, ws_bytes, s));
      B2_CUDA_OK(ctx, cudaStreamWaitEvent(s, landed.b, 0));
      B2_RETURN_NOT_OK(b2_join_u32_phased_dev(ctx, d_fk, d_y, nl, d_pk, d_x, nr, o_fk, o_y, o_x, cap, d_rows, 0, 2,
                                              d_ws, ws_bytes, s));
    } else {
      B2_RETURN_NOT_OK(b2_join_u32_dev(ctx, d_fk, d_y, nl, d_pk, d_x, nr, o_fk, o_y, o_x, cap, d_rows,
                                       0, d_ws, ws_bytes, s));
    }
    B2_CUDA_OK(ctx, cudaMemcpyAsync(&rows, d_rows, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    if (rows == ~0ull) return b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
    if ((int64_t)rows <= cap) break;
    if (attempt == 1) return b2_set_error(ctx, B2_ERR_OVERFLOW, "join", "output larger than reported");
    for (uint32_t* p : {o_fk, o_y, o_x}) {
      bufs.release(p);
      b2_dev_free(ctx, p);
    }
    cap = (int64_t)rows;
  }
  B2_CUDA_OK(ctx, cudaEventRecord(work.b, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kJoin;
  pend->d_fk = o_fk;
  pend->d_y = o_y;
  pend->d_x = o_x;
  pend->rows = rows;
  for (uint32_t* p : {o_fk, o_y, o_x}) {
    pend->dev.push_back(p);
    bufs.release(p);
  }
  ctx->pending = pend;
  *out_rows = rows;
  tm.copy_to_dev_ms = up.ms();
  tm.dev_work_ms = work.ms();
  tm.d2h_bytes = 8;
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

// Fused [filter ->] join -> aggregate over host batches: one upload, no result columns to bring back.
int b2_join_aggr_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens,
                          int64_t nl_batches, const uint32_t* const* r_ptrs, const int64_t* r_lens,
                          int64_t nr_batches, int filter_y, uint32_t y_threshold, b2_join_aggr* out,
                          b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nl_batches >= 0 && nr_batches >= 0, "negative batch count");
  B2_REQUIRE(ctx, out != nullptr, "out is null");
  B2_REQUIRE(ctx, nl_batches == 0 || (l_ptrs && l_lens), "null left batch table");
  B2_REQUIRE(ctx, nr_batches == 0 || (r_ptrs && r_lens), "null right batch table");
  int64_t nl = 0, nr = 0;
  B2_RETURN_NOT_OK(total_rows(ctx, l_lens, nl_batches, &nl));
  B2_RETURN_NOT_OK(total_rows(ctx, r_lens, nr_batches, &nr));
  b2_timings tm{};
  DevBufs bufs;
  EventPair up, work;
  B2_RETURN_NOT_OK(up.init(ctx));
  B2_RETURN_NOT_OK(work.init(ctx));
  cudaStream_t s = ctx->s_compute;
  uint32_t *d_fk, *d_y, *d_pk, *d_x;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_fk, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_y, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_pk, (size_t)nr * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_x, (size_t)nr * 4));
  B2_CUDA_OK(ctx, cudaEventRecord(up.a, s));
  B2_RETURN_NOT_OK(upload_column(ctx, d_fk, l_ptrs, l_lens, nl_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_y, l_ptrs + nl_batches, l_lens, nl_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_pk, r_ptrs, r_lens, nr_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_x, r_ptrs + nr_batches, r_lens, nr_batches, s, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(up.b, s));
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(nl, nr);
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_ws, ws_bytes));
  b2_join_aggr* d_out = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_out, sizeof(b2_join_aggr)));
  B2_CUDA_OK(ctx, cudaEventRecord(work.a, s));
  B2_RETURN_NOT_OK(b2_join_aggr_u32_dev(ctx, d_fk, d_y, nl, d_pk, d_x, nr, filter_y, y_threshold, d_out, 0,
                                        d_ws, ws_bytes, s));
  B2_CUDA_OK(ctx, cudaEventRecord(work.b, s));
  B2_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, sizeof(b2_join_aggr), cudaMemcpyDeviceToHost, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  if (out->rows == ~0ull) return b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
  tm.copy_to_dev_ms = up.ms();
  tm.dev_work_ms = work.ms();
  tm.d2h_bytes = sizeof(b2_join_aggr);
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

// ---- join over any number of payload columns per side (JoinDpu partitions every left value column,
// join_dpu.cc:127-138, and takes every right non-key column, :325-341) ------------------------------
// One payload per side travels inside the (key, payload) pair (the benchmark shape: no index vector,
// no take pass). With more, the pair carries the ROW NUMBER instead and every payload column is
// gathered afterwards with the take kernel — the reference's own scheme (selection_indices_vector +
// TakeKernel), minus the host round trips between the steps.
int b2_join_cols_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                          int nl_payloads, const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches,
                          int nr_payloads, uint64_t* out_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nl_payloads >= 0 && nl_payloads <= 64 && nr_payloads >= 0 && nr_payloads <= 64,
             "0..64 payload columns per side");
  if (nl_payloads == 1 && nr_payloads == 1)
    return b2_join_u32_host(ctx, l_ptrs, l_lens, nl_batches, r_ptrs, r_lens, nr_batches, out_rows, timings);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nl_batches >= 0 && nr_batches >= 0 && out_rows != nullptr, "bad arguments");
  B2_REQUIRE(ctx, nl_batches == 0 || (l_ptrs && l_lens), "null left batch table");
  B2_REQUIRE(ctx, nr_batches == 0 || (r_ptrs && r_lens), "null right batch table");
  int64_t nl = 0, nr = 0;
  B2_RETURN_NOT_OK(total_rows(ctx, l_lens, nl_batches, &nl));
  B2_RETURN_NOT_OK(total_rows(ctx, r_lens, nr_batches, &nr));
  B2_REQUIRE(ctx, nl < (1ll << 32) && nr < (1ll << 32), "row numbers travel as uint32: at most 2^32 - 1 rows per side");
  b2_timings tm{};
  DevBufs bufs;
  EventPair up, work;
  B2_RETURN_NOT_OK(up.init(ctx));
  B2_RETURN_NOT_OK(work.init(ctx));
  cudaStream_t s = ctx->s_compute;
  uint32_t *d_fk, *d_lid, *d_pk, *d_rid;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_fk, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_lid, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_pk, (size_t)nr * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rid, (size_t)nr * 4));
  B2_CUDA_OK(ctx, cudaEventRecord(up.a, s));
  B2_RETURN_NOT_OK(upload_column(ctx, d_fk, l_ptrs, l_lens, nl_batches, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_column(ctx, d_pk, r_ptrs, r_lens, nr_batches, s, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(up.b, s));
  B2_CUDA_OK(ctx, cudaEventRecord(work.a, s));
  B2_RETURN_NOT_OK(b2_iota_u32_dev(ctx, 0, nl, d_lid, s));
  B2_RETURN_NOT_OK(b2_iota_u32_dev(ctx, 0, nr, d_rid, s));
  int64_t cap = nl;
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(nl, nr);
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_ws, ws_bytes));
  uint64_t* d_rows = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rows, 8));
  uint32_t *o_fk = nullptr, *o_lid = nullptr, *o_rid = nullptr;
  uint64_t rows = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_fk, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_lid, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_rid, (size_t)cap * 4));
    B2_RETURN_NOT_OK(b2_join_u32_dev(ctx, d_fk, d_lid, nl, d_pk, d_rid, nr, o_fk, o_lid, o_rid, cap, d_rows, 0, d_ws,
                                     ws_bytes, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(&rows, d_rows, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    if (rows == ~0ull) return b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
    if ((int64_t)rows <= cap) break;
    if (attempt == 1) return b2_set_error(ctx, B2_ERR_OVERFLOW, "join", "output larger than reported");
    for (uint32_t* p : {o_fk, o_lid, o_rid}) {
      bufs.release(p);
      b2_dev_free(ctx, p);
    }
    cap = (int64_t)rows;
  }
  // the inputs of the join are no longer needed; the payload columns take their place one at a time
  for (void* p : {(void*)d_fk, (void*)d_pk, (void*)d_lid, (void*)d_rid, d_ws}) {
    bufs.release(p);
    b2_dev_free(ctx, p);
  }
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kJoinCols;
  pend->rows = rows;
  pend->d_fk = o_fk;
  pend->dev.push_back(o_fk);
  bufs.release(o_fk);
  ctx->pending = pend;
  uint32_t* d_col = nullptr;
  const int64_t side_rows[2] = {nl, nr};
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_col, (size_t)std::max(nl, nr) * 4));
  for (int side = 0; side < 2; ++side) {
    const int np = side == 0 ? nl_payloads : nr_payloads;
    const uint32_t* const* ptrs = side == 0 ? l_ptrs : r_ptrs;
    const int64_t* lens = side == 0 ? l_lens : r_lens;
    const int64_t nb = side == 0 ? nl_batches : nr_batches;
    const uint32_t* ids = side == 0 ? o_lid : o_rid;
    for (int c = 0; c < np; ++c) {
      uint32_t* d_out = nullptr;
      B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_out, (size_t)std::max<uint64_t>(rows, 1) * 4));
      pend->dev.push_back(d_out);
      pend->d_cols.push_back(d_out);
      B2_RETURN_NOT_OK(upload_column(ctx, d_col, ptrs + (size_t)(1 + c) * nb, lens, nb, s, &tm.h2d_bytes));
      if (rows > 0) {  // one "batch" = the whole side: the row numbers are global
        b2_trace_scope tr(ctx, B2_PHASE_TAKE, s);
        B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, d_col, side_rows[side], ids, (int64_t)rows, 1, d_out, s));
      }
    }
  }
  B2_CUDA_OK(ctx, cudaEventRecord(work.b, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  pend->ncols = 1 + nl_payloads + nr_payloads;
  *out_rows = rows;
  tm.copy_to_dev_ms = up.ms();
  tm.dev_work_ms = work.ms();  // includes the payload uploads interleaved with the gathers
  tm.d2h_bytes = 8;
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_join_cols_fetch_host(b2_ctx* ctx, uint32_t* const* out_cols, int ncols, int64_t capacity_rows,
                            b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  b2_pending* pend = ctx->pending;
  B2_REQUIRE(ctx, out_cols != nullptr || ncols == 0, "out_cols is null");
  if (pend && pend->kind == b2_pending::kJoin && ncols == 3)  // the one-payload fast path ran
    return b2_join_fetch_host(ctx, out_cols[0], out_cols[1], out_cols[2], capacity_rows, timings);
  if (!pend || pend->kind != b2_pending::kJoinCols)
    return b2_set_error(ctx, B2_ERR_INVALID, "b2_join_cols_fetch_host", "no pending join result");
  B2_REQUIRE(ctx, ncols == pend->ncols, "column count differs from the run");
  if ((uint64_t)capacity_rows < pend->rows)
    return b2_set_error(ctx, B2_ERR_OVERFLOW, "b2_join_cols_fetch_host", "capacity_rows < result rows");
  b2_timings tm{};
  if (pend->rows > 0) {
    EventPair ev;
    B2_RETURN_NOT_OK(ev.init(ctx));
    cudaStream_t s = ctx->s_copy_out;
    const size_t bytes = (size_t)pend->rows * 4;
    B2_CUDA_OK(ctx, cudaEventRecord(ev.a, s));
    for (int c = 0; c < ncols; ++c) {
      B2_REQUIRE(ctx, out_cols[c] != nullptr, "null output column");
      const uint32_t* src = c == 0 ? pend->d_fk : pend->d_cols[(size_t)c - 1];
      B2_CUDA_OK(ctx, cudaMemcpyAsync(out_cols[c], src, bytes, cudaMemcpyDeviceToHost, s));
      tm.d2h_bytes += (int64_t)bytes;
    }
    B2_CUDA_OK(ctx, cudaEventRecord(ev.b, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    tm.copy_from_dev_ms = ev.ms();
  }
  tm.total_ms = ms_since(t0);
  if (timings) *timings = tm;
  return B2_OK;
}

// ---- join over a typed table: 32- or 64-bit keys, any number of 32- or 64-bit payload columns ---------
// The reference's table can hold 64-bit keys (HT_64BIT_KEYS, dpu/shared/hashtable/hashtable.h:14-18) and
// JoinDpu carries every non-key column (join_dpu.cc:127-138,325-341). The radix passes and the probe
// kernel move 8-byte (key, payload) pairs; wider rows go through them by reference:
//   * the pair's payload is the ROW NUMBER, every output column is gathered afterwards with the take
//     kernels (32- and 64-bit);
//   * a 64-bit key is folded to 32 bits for partitioning and probing — equal keys have equal folds, so
//     the candidate (left row, right row) pairs are a superset of the result — and every candidate is
//     verified against the full 64-bit keys; the survivors are compacted with the filter kernel.
namespace {

__global__ void __launch_bounds__(256)
fold_u64_kernel(const uint64_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const uint64_t k = in[i];
    out[i] = (uint32_t)k ^ ((uint32_t)(k >> 32) * 0x9E3779B1u);
  }
}
// sel[i] = i when candidate pair i joins two rows with the SAME 64-bit key, else 0xffffffff
__global__ void __launch_bounds__(256)
verify_u64_kernel(const uint64_t* __restrict__ lkeys, const uint64_t* __restrict__ rkeys,
                  const uint32_t* __restrict__ lid, const uint32_t* __restrict__ rid, int64_t n,
                  uint32_t* __restrict__ sel) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    sel[i] = lkeys[lid[i]] == rkeys[rid[i]] ? (uint32_t)i : 0xffffffffu;
}

int upload_wide(b2_ctx* ctx, void* d_col, const void* const* ptrs, const int64_t* lens, int64_t nbatches, int bytes,
                cudaStream_t s, int64_t* total) {
  char* dst = static_cast<char*>(d_col);
  for (int64_t b = 0; b < nbatches; ++b) {
    if (lens[b] == 0) continue;
    B2_REQUIRE(ctx, ptrs[b] != nullptr, "null batch pointer");
    B2_CUDA_OK(ctx, cudaMemcpyAsync(dst, ptrs[b], (size_t)lens[b] * bytes, cudaMemcpyHostToDevice, s));
    dst += (size_t)lens[b] * bytes;
    *total += lens[b] * bytes;
  }
  return B2_OK;
}

int gather_col(b2_ctx* ctx, const void* d_col, int64_t col_rows, int bytes, const uint32_t* ids, int64_t rows,
               void* d_out, cudaStream_t s) {
  if (rows == 0) return B2_OK;
  b2_trace_scope tr(ctx, B2_PHASE_TAKE, s);
  if (bytes == 8) return b2_take_64_dev(ctx, d_col, col_rows, ids, rows, 1, d_out, s);
  return b2_take_u32_dev(ctx, static_cast<const uint32_t*>(d_col), col_rows, ids, rows, 1,
                         static_cast<uint32_t*>(d_out), s);
}

}  // namespace

int b2_join_table_host(b2_ctx* ctx, const void* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                       const int* l_col_bytes, int nl_cols, const void* const* r_ptrs, const int64_t* r_lens,
                       int64_t nr_batches, const int* r_col_bytes, int nr_cols, uint64_t* out_rows,
                       b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  const int64_t launches0 = ctx->launches;
  b2_pending_free(ctx);
  B2_RETURN_NOT_OK(ensure_streams(ctx));
  B2_REQUIRE(ctx, nl_cols >= 1 && nl_cols <= 65 && nr_cols >= 1 && nr_cols <= 65, "1 key + 0..64 payload columns per side");
  B2_REQUIRE(ctx, l_col_bytes && r_col_bytes && out_rows, "null argument");
  B2_REQUIRE(ctx, nl_batches >= 0 && nr_batches >= 0, "negative batch count");
  B2_REQUIRE(ctx, nl_batches == 0 || (l_ptrs && l_lens), "null left batch table");
  B2_REQUIRE(ctx, nr_batches == 0 || (r_ptrs && r_lens), "null right batch table");
  for (int c = 0; c < nl_cols; ++c) B2_REQUIRE(ctx, l_col_bytes[c] == 4 || l_col_bytes[c] == 8, "columns are 4 or 8 bytes wide");
  for (int c = 0; c < nr_cols; ++c) B2_REQUIRE(ctx, r_col_bytes[c] == 4 || r_col_bytes[c] == 8, "columns are 4 or 8 bytes wide");
  B2_REQUIRE(ctx, l_col_bytes[0] == r_col_bytes[0], "both key columns must have the same width");
  const int kb = l_col_bytes[0];
  int64_t nl = 0, nr = 0;
  B2_RETURN_NOT_OK(total_rows(ctx, l_lens, nl_batches, &nl));
  B2_RETURN_NOT_OK(total_rows(ctx, r_lens, nr_batches, &nr));
  B2_REQUIRE(ctx, nl < (1ll << 32) - 1 && nr < (1ll << 32) - 1, "row numbers travel as uint32");
  b2_timings tm{};
  DevBufs bufs;
  EventPair up, work;
  B2_RETURN_NOT_OK(up.init(ctx));
  B2_RETURN_NOT_OK(work.init(ctx));
  cudaStream_t s = ctx->s_compute;
  const int64_t grid_cap = (int64_t)ctx->sm_count * 16;
  auto grid_for = [&](int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, grid_cap)); };
  void *d_lkey = nullptr, *d_rkey = nullptr;
  uint32_t *d_lk32 = nullptr, *d_rk32 = nullptr, *d_lid = nullptr, *d_rid = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_lkey, (size_t)nl * kb));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_rkey, (size_t)nr * kb));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_lid, (size_t)nl * 4));
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rid, (size_t)nr * 4));
  B2_CUDA_OK(ctx, cudaEventRecord(up.a, s));
  B2_RETURN_NOT_OK(upload_wide(ctx, d_lkey, l_ptrs, l_lens, nl_batches, kb, s, &tm.h2d_bytes));
  B2_RETURN_NOT_OK(upload_wide(ctx, d_rkey, r_ptrs, r_lens, nr_batches, kb, s, &tm.h2d_bytes));
  B2_CUDA_OK(ctx, cudaEventRecord(up.b, s));
  B2_CUDA_OK(ctx, cudaEventRecord(work.a, s));
  if (kb == 8) {
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_lk32, (size_t)nl * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rk32, (size_t)nr * 4));
    if (nl > 0) {
      fold_u64_kernel<<<grid_for(nl), 256, 0, s>>>(static_cast<const uint64_t*>(d_lkey), nl, d_lk32);
      B2_LAUNCH_CHECK(ctx, "fold_u64_kernel");
    }
    if (nr > 0) {
      fold_u64_kernel<<<grid_for(nr), 256, 0, s>>>(static_cast<const uint64_t*>(d_rkey), nr, d_rk32);
      B2_LAUNCH_CHECK(ctx, "fold_u64_kernel");
    }
  } else {
    d_lk32 = static_cast<uint32_t*>(d_lkey);
    d_rk32 = static_cast<uint32_t*>(d_rkey);
  }
  B2_RETURN_NOT_OK(b2_iota_u32_dev(ctx, 0, nl, d_lid, s));
  B2_RETURN_NOT_OK(b2_iota_u32_dev(ctx, 0, nr, d_rid, s));
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(nl, nr);
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_ws, ws_bytes));
  uint64_t* d_rows = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_rows, 8));
  int64_t cap = nl;
  uint32_t *o_k = nullptr, *o_lid = nullptr, *o_rid = nullptr;
  uint64_t cand = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_k, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_lid, (size_t)cap * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&o_rid, (size_t)cap * 4));
    B2_RETURN_NOT_OK(b2_join_u32_dev(ctx, d_lk32, d_lid, nl, d_rk32, d_rid, nr, o_k, o_lid, o_rid, cap, d_rows, 0, d_ws,
                                     ws_bytes, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(&cand, d_rows, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    if (cand == ~0ull) return b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
    if ((int64_t)cand <= cap) break;
    if (attempt == 1) return b2_set_error(ctx, B2_ERR_OVERFLOW, "join", "output larger than reported");
    for (uint32_t* p : {o_k, o_lid, o_rid}) {
      bufs.release(p);
      b2_dev_free(ctx, p);
    }
    cap = (int64_t)cand;
  }
  B2_REQUIRE(ctx, cand < 0xffffffffull, "more than 2^32 - 2 candidate pairs");
  bufs.release(d_ws);
  b2_dev_free(ctx, d_ws);
  int64_t rows = (int64_t)cand;
  const uint32_t *lids = o_lid, *rids = o_rid;
  if (kb == 8 && cand > 0) {  // folds collide: keep the candidates whose full keys are equal
    uint32_t *d_sel = nullptr, *d_surv = nullptr, *d_l2 = nullptr, *d_r2 = nullptr;
    int64_t *d_end = nullptr, *d_total = nullptr;
    void* d_fws = nullptr;
    const size_t fws = b2_filter_ws_bytes(1, (int64_t)cand);
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_sel, (size_t)cand * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_surv, (size_t)cand * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_end, 8));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_total, 8));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_fws, fws));
    verify_u64_kernel<<<grid_for((int64_t)cand), 256, 0, s>>>(static_cast<const uint64_t*>(d_lkey),
                                                              static_cast<const uint64_t*>(d_rkey), o_lid, o_rid,
                                                              (int64_t)cand, d_sel);
    B2_LAUNCH_CHECK(ctx, "verify_u64_kernel");
    B2_RETURN_NOT_OK(b2_filter_lt_u32_dev(ctx, d_sel, 1, (int64_t)cand, 0xffffffffu, d_surv, d_end, d_total, nullptr,
                                          d_fws, fws, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(&rows, d_total, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_l2, (size_t)std::max<int64_t>(rows, 1) * 4));
    B2_RETURN_NOT_OK(bufs.alloc(ctx, (void**)&d_r2, (size_t)std::max<int64_t>(rows, 1) * 4));
    if (rows > 0) {
      B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, o_lid, (int64_t)cand, d_surv, rows, 1, d_l2, s));
      B2_RETURN_NOT_OK(b2_take_u32_dev(ctx, o_rid, (int64_t)cand, d_surv, rows, 1, d_r2, s));
    }
    lids = d_l2;
    rids = d_r2;
  }
  // result columns: the key, the left payloads, the right payloads — each gathered by row number
  b2_pending* pend = new b2_pending();
  pend->kind = b2_pending::kJoinTable;
  pend->rows = (uint64_t)rows;
  ctx->pending = pend;
  auto add_out = [&](int bytes, void** out) -> int {
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, out, (size_t)std::max<int64_t>(rows, 1) * bytes));
    pend->dev.push_back(*out);
    pend->t_cols.push_back(*out);
    pend->t_bytes.push_back(bytes);
    return B2_OK;
  };
  void* d_out = nullptr;
  B2_RETURN_NOT_OK(add_out(kb, &d_out));
  B2_RETURN_NOT_OK(gather_col(ctx, d_lkey, nl, kb, lids, rows, d_out, s));
  void* d_col = nullptr;
  B2_RETURN_NOT_OK(bufs.alloc(ctx, &d_col, (size_t)std::max(nl, nr) * 8));
  for (int side = 0; side < 2; ++side) {
    const int nc = side == 0 ? nl_cols : nr_cols;
    const int* cb = side == 0 ? l_col_bytes : r_col_bytes;
    const void* const* ptrs = side == 0 ? l_ptrs : r_ptrs;
    const int64_t* lens = side == 0 ? l_lens : r_lens;
    const int64_t nb = side == 0 ? nl_batches : nr_batches;
    const int64_t n = side == 0 ? nl : nr;
    for (int c = 1; c < nc; ++c) {
      B2_RETURN_NOT_OK(add_out(cb[c], &d_out));
      B2_RETURN_NOT_OK(upload_wide(ctx, d_col, ptrs + (size_t)c * nb, lens, nb, cb[c], s, &tm.h2d_bytes));
      B2_RETURN_NOT_OK(gather_col(ctx, d_col, n, cb[c], side == 0 ? lids : rids, rows, d_out, s));
    }
  }
  B2_CUDA_OK(ctx, cudaEventRecord(work.b, s));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  pend->ncols = (int)pend->t_cols.size();
  *out_rows = (uint64_t)rows;
  tm.copy_to_dev_ms = up.ms();
  tm.dev_work_ms = work.ms();
  tm.d2h_bytes = 8;
  tm.total_ms = ms_since(t0);
  tm.kernel_launches = (int32_t)(ctx->launches - launches0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_join_table_fetch_host(b2_ctx* ctx, void* const* out_cols, int ncols, int64_t capacity_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  b2_pending* pend = ctx->pending;
  if (!pend || pend->kind != b2_pending::kJoinTable)
    return b2_set_error(ctx, B2_ERR_INVALID, "b2_join_table_fetch_host", "no pending join result");
  B2_REQUIRE(ctx, ncols == pend->ncols && (out_cols || ncols == 0), "column count differs from the run");
  if ((uint64_t)capacity_rows < pend->rows)
    return b2_set_error(ctx, B2_ERR_OVERFLOW, "b2_join_table_fetch_host", "capacity_rows < result rows");
  b2_timings tm{};
  if (pend->rows > 0) {
    EventPair ev;
    B2_RETURN_NOT_OK(ev.init(ctx));
    cudaStream_t s = ctx->s_copy_out;
    B2_CUDA_OK(ctx, cudaEventRecord(ev.a, s));
    for (int c = 0; c < ncols; ++c) {
      B2_REQUIRE(ctx, out_cols[c] != nullptr, "null output column");
      const size_t bytes = (size_t)pend->rows * (size_t)pend->t_bytes[(size_t)c];
      B2_CUDA_OK(ctx, cudaMemcpyAsync(out_cols[c], pend->t_cols[(size_t)c], bytes, cudaMemcpyDeviceToHost, s));
      tm.d2h_bytes += (int64_t)bytes;
    }
    B2_CUDA_OK(ctx, cudaEventRecord(ev.b, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    tm.copy_from_dev_ms = ev.ms();
  }
  tm.total_ms = ms_since(t0);
  if (timings) *timings = tm;
  return B2_OK;
}

int b2_join_fetch_host(b2_ctx* ctx, uint32_t* out_fk, uint32_t* out_y, uint32_t* out_x,
                       int64_t capacity_rows, b2_timings* timings) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  const auto t0 = Clock::now();
  b2_pending* pend = ctx->pending;
  if (!pend || pend->kind != b2_pending::kJoin)
    return b2_set_error(ctx, B2_ERR_INVALID, "b2_join_fetch_host", "no pending join result");
  if ((uint64_t)capacity_rows < pend->rows)
    return b2_set_error(ctx, B2_ERR_OVERFLOW, "b2_join_fetch_host", "capacity_rows < result rows");
  b2_timings tm{};
  if (pend->rows > 0) {
    B2_REQUIRE(ctx, out_fk && out_y && out_x, "null output column");
    EventPair ev;
    B2_RETURN_NOT_OK(ev.init(ctx));
    cudaStream_t s = ctx->s_copy_out;
    const size_t bytes = (size_t)pend->rows * 4;
    B2_CUDA_OK(ctx, cudaEventRecord(ev.a, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_fk, pend->d_fk, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_y, pend->d_y, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_x, pend->d_x, bytes, cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(ctx, cudaEventRecord(ev.b, s));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
    tm.copy_from_dev_ms = ev.ms();
    tm.d2h_bytes = (int64_t)bytes * 3;
  }
  tm.total_ms = ms_since(t0);
  if (timings) *timings = tm;
  return B2_OK;
}

}  // extern "C"
