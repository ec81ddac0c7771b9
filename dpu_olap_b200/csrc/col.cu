// col.cu — device-resident columns (b2_col) and their Arrow C Device Data Interface view.
//
// The reference moves every operator's input host -> DPU and its output DPU -> host
// (arrow_copy_to_dpus / arrow_copy_from_dpus*, host/dpuext/arrow_utils.cc:47-73,147-266) and notes
// itself that results should not have to be copied (arrow_utils.h:28-29). A b2_col is one packed
// uint32 column in HBM plus its batch boundaries; operators take and return b2_cols, so a chain such
// as filter -> take -> sum or join -> sum never crosses PCIe, and a column can be handed to (or
// taken from) any Arrow consumer on the same GPU as an ArrowDeviceArray (device_type CUDA, buffers[1]
// = the device pointer, sync_event = a cudaEvent_t recorded after the producing kernels) with no copy.
#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"

struct b2_col {
  b2_ctx* ctx = nullptr;
  uint32_t* d = nullptr;            // packed column
  int64_t rows = 0;
  std::vector<int64_t> off;         // batch boundaries in rows (nbatches + 1); batches are back to back
  std::atomic<int> refs{1};         // the handle + every exported ArrowDeviceArray
  cudaEvent_t ready = nullptr;      // recorded on the ctx's compute stream after the producer
  bool pooled = false;              // d came from the ctx's recycling allocator
  ArrowDeviceArray* imported = nullptr;  // the column VIEWS an imported array: released with the column
};

namespace {

int ensure_stream(b2_ctx* ctx) {
  if (!ctx->s_compute) B2_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  return B2_OK;
}

void col_unref(b2_col* c) {
  if (!c || c->refs.fetch_sub(1) != 1) return;
  b2_device_scope sc(c->ctx);
  if (c->ready) {
    cudaEventSynchronize(c->ready);
    cudaEventDestroy(c->ready);
  }
  if (c->imported) {
    if (c->imported->array.release) c->imported->array.release(&c->imported->array);
    delete c->imported;
  } else if (c->d) {
    if (c->pooled) b2_dev_free(c->ctx, c->d);
    else cudaFree(c->d);
  }
  delete c;
}

int col_new(b2_ctx* ctx, int64_t rows, b2_col** out) {
  b2_col* c = new b2_col();
  c->ctx = ctx;
  c->rows = rows;
  c->pooled = true;
  int rc = b2_dev_alloc(ctx, (void**)&c->d, (size_t)std::max<int64_t>(rows, 1) * 4);
  if (rc == B2_OK && cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess)
    rc = b2_set_error(ctx, B2_ERR_CUDA, "cudaEventCreate", nullptr);
  if (rc != B2_OK) {
    col_unref(c);
    return rc;
  }
  *out = c;
  return B2_OK;
}

int col_done(b2_col* c) {  // the producer has been enqueued: mark the point consumers must wait for
  B2_CUDA_OK(c->ctx, cudaEventRecord(c->ready, c->ctx->s_compute));
  return B2_OK;
}

// consumers run on the same compute stream, so stream order already covers library-produced
// columns; imported columns carry a foreign event
int col_wait(b2_ctx* ctx, const b2_col* c) {
  if (c && c->imported && c->imported->sync_event)
    B2_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_compute, *static_cast<cudaEvent_t*>(c->imported->sync_event), 0));
  return B2_OK;
}

bool uniform(const b2_col* c, int64_t* batch_len) {
  const int64_t nb = (int64_t)c->off.size() - 1;
  if (nb <= 0) {
    *batch_len = 0;
    return true;
  }
  const int64_t l = c->off[1] - c->off[0];
  for (int64_t b = 1; b < nb; ++b)
    if (c->off[(size_t)b + 1] - c->off[(size_t)b] != l) return false;
  *batch_len = l;
  return true;
}

struct ExportPrivate {
  b2_col* col;
  const void* buffers[2];
  cudaEvent_t event;
};

void export_release(ArrowArray* a) {
  if (!a || !a->release) return;
  ExportPrivate* p = static_cast<ExportPrivate*>(a->private_data);
  col_unref(p->col);
  delete p;
  a->release = nullptr;
}

}  // namespace

extern "C" {

int b2_col_upload_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const int64_t* batch_lens, int64_t nbatches,
                       b2_col** out) {
  if (!ctx || !out) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  *out = nullptr;
  B2_REQUIRE(ctx, nbatches >= 0 && (nbatches == 0 || (batch_ptrs && batch_lens)), "bad batch table");
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  int64_t rows = 0;
  for (int64_t b = 0; b < nbatches; ++b) {
    B2_REQUIRE(ctx, batch_lens[b] >= 0, "negative batch length");
    rows += batch_lens[b];
  }
  b2_col* c = nullptr;
  B2_RETURN_NOT_OK(col_new(ctx, rows, &c));
  c->off.assign((size_t)nbatches + 1, 0);
  int64_t o = 0, b = 0;
  while (b < nbatches) {  // host-adjacent batches travel as one copy
    int64_t e = b + 1, n = batch_lens[b];
    while (e < nbatches && batch_ptrs[e] == batch_ptrs[e - 1] + batch_lens[e - 1]) n += batch_lens[e++];
    if (n > 0 && cudaMemcpyAsync(c->d + o, batch_ptrs[b], (size_t)n * 4, cudaMemcpyHostToDevice, ctx->s_compute) !=
                     cudaSuccess) {
      col_unref(c);
      return b2_set_error(ctx, B2_ERR_CUDA, "cudaMemcpyAsync", cudaGetErrorString(cudaGetLastError()));
    }
    for (int64_t k = b; k < e; ++k) {
      c->off[(size_t)k + 1] = c->off[(size_t)k] + batch_lens[k];
    }
    o += n;
    b = e;
  }
  const int rc = col_done(c);
  if (rc != B2_OK) {
    col_unref(c);
    return rc;
  }
  *out = c;
  return B2_OK;
}

int b2_col_free(b2_col* col) {
  col_unref(col);
  return B2_OK;
}
int64_t b2_col_rows(const b2_col* col) { return col ? col->rows : 0; }
int64_t b2_col_nbatches(const b2_col* col) { return col ? (int64_t)col->off.size() - 1 : 0; }
const uint32_t* b2_col_device_ptr(const b2_col* col) { return col ? col->d : nullptr; }
int b2_col_batch_offsets(const b2_col* col, int64_t* out, int64_t capacity) {
  if (!col || !out || capacity < (int64_t)col->off.size()) return B2_ERR_INVALID;
  std::copy(col->off.begin(), col->off.end(), out);
  return B2_OK;
}

int b2_col_download_host(b2_ctx* ctx, const b2_col* col, uint32_t* const* out_ptrs, int64_t nbatches) {
  if (!ctx || !col) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches == (int64_t)col->off.size() - 1, "batch count differs from the column's");
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  B2_RETURN_NOT_OK(col_wait(ctx, col));
  for (int64_t b = 0; b < nbatches; ++b) {
    const int64_t r0 = col->off[(size_t)b], n = col->off[(size_t)b + 1] - r0;
    if (n == 0) continue;
    B2_REQUIRE(ctx, out_ptrs && out_ptrs[b], "null output pointer for a non-empty batch");
    B2_CUDA_OK(ctx, cudaMemcpyAsync(out_ptrs[b], col->d + r0, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->s_compute));
  }
  B2_CUDA_OK(ctx, cudaStreamSynchronize(ctx->s_compute));
  return B2_OK;
}

// ---- Arrow C Device Data Interface -------------------------------------------------------------------
int b2_col_export(b2_col* col, ArrowDeviceArray* out) {
  if (!col || !out) return B2_ERR_INVALID;
  ExportPrivate* p = new ExportPrivate();
  p->col = col;
  p->buffers[0] = nullptr;  // no validity bitmap: the column is non-null
  p->buffers[1] = col->d;
  p->event = col->ready;
  col->refs.fetch_add(1);  // the exported array keeps the column (and its device memory) alive
  ArrowArray& a = out->array;
  a.length = col->rows;
  a.null_count = 0;
  a.offset = 0;
  a.n_buffers = 2;
  a.n_children = 0;
  a.buffers = p->buffers;
  a.children = nullptr;
  a.dictionary = nullptr;
  a.release = export_release;
  a.private_data = p;
  out->device_id = col->ctx->device;
  out->device_type = ARROW_DEVICE_CUDA;
  out->sync_event = &p->event;  // cudaEvent_t*: recorded after the kernels that produced the column
  out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
  return B2_OK;
}

int b2_col_import(b2_ctx* ctx, ArrowDeviceArray* in, const int64_t* batch_lens, int64_t nbatches, b2_col** out) {
  if (!ctx || !in || !out) return B2_ERR_INVALID;
  *out = nullptr;
  B2_REQUIRE(ctx, in->array.release != nullptr, "the array has been released");
  B2_REQUIRE(ctx, in->device_type == ARROW_DEVICE_CUDA && in->device_id == ctx->device,
             "the array must live on this context's CUDA device");
  B2_REQUIRE(ctx, in->array.n_buffers == 2 && in->array.n_children == 0 && in->array.dictionary == nullptr,
             "a primitive uint32 array has two buffers and no children");
  B2_REQUIRE(ctx, in->array.null_count == 0 || in->array.buffers[0] == nullptr, "nullable columns are not taken here");
  int64_t rows = 0;
  for (int64_t b = 0; b < nbatches; ++b) rows += batch_lens[b];
  if (nbatches == 0) rows = in->array.length;
  B2_REQUIRE(ctx, rows == in->array.length, "batch lengths do not add up to the array's length");
  b2_col* c = new b2_col();
  c->ctx = ctx;
  c->rows = rows;
  c->d = const_cast<uint32_t*>(static_cast<const uint32_t*>(in->array.buffers[1])) + in->array.offset;
  c->off.assign(1, 0);
  if (nbatches == 0) c->off.push_back(rows);
  for (int64_t b = 0; b < nbatches; ++b) c->off.push_back(c->off.back() + batch_lens[b]);
  c->imported = new ArrowDeviceArray(*in);  // ownership moves: the producer's release runs when the column dies
  in->array.release = nullptr;
  b2_device_scope dev_scope(ctx);
  if (cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess) {
    col_unref(c);
    return b2_set_error(ctx, B2_ERR_CUDA, "cudaEventCreate", nullptr);
  }
  *out = c;
  return B2_OK;
}

// ---- operators on device columns: results stay in HBM ------------------------------------------------
int b2_sum_u32_col(b2_ctx* ctx, const b2_col* col, uint64_t* sum) {
  if (!ctx || !col || !sum) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  B2_RETURN_NOT_OK(col_wait(ctx, col));
  uint64_t* d_sum = nullptr;
  B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_sum, 256));
  int rc = b2_sum_u32_dev(ctx, col->d, col->rows, d_sum, ctx->s_compute);
  if (rc == B2_OK && (cudaMemcpyAsync(sum, d_sum, 8, cudaMemcpyDeviceToHost, ctx->s_compute) != cudaSuccess ||
                      cudaStreamSynchronize(ctx->s_compute) != cudaSuccess))
    rc = b2_set_error(ctx, B2_ERR_CUDA, "sum read-back", cudaGetErrorString(cudaGetLastError()));
  b2_dev_free(ctx, d_sum);
  return rc;
}

int b2_filter_lt_u32_col(b2_ctx* ctx, const b2_col* col, uint32_t threshold, b2_col** out) {
  if (!ctx || !col || !out) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  *out = nullptr;
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  B2_RETURN_NOT_OK(col_wait(ctx, col));
  cudaStream_t s = ctx->s_compute;
  const int64_t nb = (int64_t)col->off.size() - 1;
  int64_t batch_len = 0;
  const bool uni = uniform(col, &batch_len);
  b2_col* res = nullptr;
  B2_RETURN_NOT_OK(col_new(ctx, col->rows, &res));  // worst case: every row is selected
  struct Guard {
    b2_ctx* ctx;
    std::vector<void*> bufs;
    b2_col* res;
    ~Guard() {
      for (void* p : bufs) b2_dev_free(ctx, p);
      if (res) col_unref(res);
    }
  } g{ctx, {}, res};
  int64_t* d_end = nullptr;
  B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_end, (size_t)std::max<int64_t>(nb, 1) * 8));
  g.bufs.push_back(d_end);
  void* d_ws = nullptr;
  size_t ws_bytes = 0;
  int64_t* d_off = nullptr;
  if (uni) {
    ws_bytes = b2_filter_ws_bytes(nb, batch_len);
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &d_ws, ws_bytes));
    g.bufs.push_back(d_ws);
    B2_RETURN_NOT_OK(b2_filter_lt_u32_dev(ctx, col->d, nb, batch_len, threshold, res->d, d_end, nullptr, nullptr, d_ws,
                                          ws_bytes, s));
  } else {
    ws_bytes = b2_filter_ragged_ws_bytes(col->off.data(), nb);
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &d_ws, ws_bytes));
    g.bufs.push_back(d_ws);
    B2_RETURN_NOT_OK(b2_dev_alloc(ctx, (void**)&d_off, (size_t)(nb + 1) * 8));
    g.bufs.push_back(d_off);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(d_off, col->off.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    B2_RETURN_NOT_OK(b2_filter_lt_u32_ragged_dev(ctx, col->d, col->off.data(), d_off, nb, threshold, res->d, d_end,
                                                 nullptr, nullptr, d_ws, ws_bytes, s));
  }
  // the chunk boundaries of the result (one chunk per input batch, filter_dpu.cc:89-96) are host-side
  // metadata of the column: 8 bytes per batch come back, the rows stay in HBM
  std::vector<int64_t> ends((size_t)std::max<int64_t>(nb, 1));
  if (nb > 0) B2_CUDA_OK(ctx, cudaMemcpyAsync(ends.data(), d_end, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
  B2_RETURN_NOT_OK(col_done(res));
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  res->off.assign(1, 0);
  for (int64_t b = 0; b < nb; ++b) res->off.push_back(ends[(size_t)b]);
  res->rows = nb > 0 ? ends[(size_t)nb - 1] : 0;
  g.res = nullptr;
  *out = res;
  return B2_OK;
}

int b2_take_u32_col(b2_ctx* ctx, const b2_col* values, const b2_col* indices, b2_col** out) {
  if (!ctx || !values || !indices || !out) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  *out = nullptr;
  B2_REQUIRE(ctx, values->off.size() == indices->off.size(), "values and indices must have the same number of batches");
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  B2_RETURN_NOT_OK(col_wait(ctx, values));
  B2_RETURN_NOT_OK(col_wait(ctx, indices));
  cudaStream_t s = ctx->s_compute;
  const int64_t nb = (int64_t)values->off.size() - 1;
  b2_col* res = nullptr;
  B2_RETURN_NOT_OK(col_new(ctx, indices->rows, &res));
  res->off = indices->off;
  int64_t vlen = 0, ilen = 0;
  int rc = B2_OK;
  if (uniform(values, &vlen) && uniform(indices, &ilen)) {
    rc = b2_take_u32_dev(ctx, values->d, vlen, indices->d, ilen, nb, res->d, s);
  } else {
    int64_t* d_tab = nullptr;
    rc = b2_dev_alloc(ctx, (void**)&d_tab, (size_t)(nb + 1) * 16);
    if (rc == B2_OK) {
      if (cudaMemcpyAsync(d_tab, values->off.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
          cudaMemcpyAsync(d_tab + nb + 1, indices->off.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s) !=
              cudaSuccess)
        rc = b2_set_error(ctx, B2_ERR_CUDA, "offset tables", cudaGetErrorString(cudaGetLastError()));
      if (rc == B2_OK)
        rc = b2_take_u32_ragged_dev(ctx, values->d, d_tab, indices->d, d_tab + nb + 1, nb, 0, indices->rows, res->d, s);
      cudaStreamSynchronize(s);  // the tables must outlive the kernel
      b2_dev_free(ctx, d_tab);
    }
  }
  if (rc == B2_OK) rc = col_done(res);
  if (rc != B2_OK) {
    col_unref(res);
    return rc;
  }
  *out = res;
  return B2_OK;
}

int b2_join_u32_col(b2_ctx* ctx, const b2_col* fk, const b2_col* y, const b2_col* pk, const b2_col* x, b2_col** out_fk,
                    b2_col** out_y, b2_col** out_x) {
  if (!ctx || !fk || !y || !pk || !x || !out_fk || !out_y || !out_x) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  *out_fk = *out_y = *out_x = nullptr;
  B2_REQUIRE(ctx, fk->rows == y->rows && pk->rows == x->rows, "key and payload columns of a side have equal lengths");
  B2_RETURN_NOT_OK(ensure_stream(ctx));
  for (const b2_col* c : {fk, y, pk, x}) B2_RETURN_NOT_OK(col_wait(ctx, c));
  cudaStream_t s = ctx->s_compute;
  const int64_t nl = fk->rows, nr = pk->rows;
  void* d_ws = nullptr;
  const size_t ws_bytes = b2_join_ws_bytes(nl, nr);
  B2_RETURN_NOT_OK(b2_dev_alloc(ctx, &d_ws, ws_bytes));
  uint64_t* d_rows = nullptr;
  int rc = b2_dev_alloc(ctx, (void**)&d_rows, 256);
  b2_col* o[3] = {nullptr, nullptr, nullptr};
  int64_t cap = nl;  // PK-FK joins; duplicate build keys re-run with the count they reported
  uint64_t rows = 0;
  for (int attempt = 0; attempt < 2 && rc == B2_OK; ++attempt) {
    for (int c = 0; c < 3 && rc == B2_OK; ++c) rc = col_new(ctx, cap, &o[c]);
    if (rc == B2_OK)
      rc = b2_join_u32_dev(ctx, fk->d, y->d, nl, pk->d, x->d, nr, o[0]->d, o[1]->d, o[2]->d, cap, d_rows, 0, d_ws,
                           ws_bytes, s);
    if (rc == B2_OK && (cudaMemcpyAsync(&rows, d_rows, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                        cudaStreamSynchronize(s) != cudaSuccess))
      rc = b2_set_error(ctx, B2_ERR_CUDA, "join row count", cudaGetErrorString(cudaGetLastError()));
    if (rc == B2_OK && rows == ~0ull) rc = b2_set_error(ctx, B2_ERR_WORKSPACE, "join", "slice overflow");
    if (rc != B2_OK || (int64_t)rows <= cap) break;
    for (int c = 0; c < 3; ++c) {
      col_unref(o[c]);
      o[c] = nullptr;
    }
    cap = (int64_t)rows;
    if (attempt == 1) rc = b2_set_error(ctx, B2_ERR_OVERFLOW, "join", "output larger than reported");
  }
  b2_dev_free(ctx, d_ws);
  if (d_rows) b2_dev_free(ctx, d_rows);
  for (int c = 0; c < 3 && rc == B2_OK; ++c) {
    o[c]->rows = (int64_t)rows;
    o[c]->off = {0, (int64_t)rows};  // one chunk: the join's row order is unspecified anyway
    rc = col_done(o[c]);
  }
  if (rc != B2_OK) {
    for (int c = 0; c < 3; ++c)
      if (o[c]) col_unref(o[c]);
    return rc;
  }
  *out_fk = o[0];
  *out_y = o[1];
  *out_x = o[2];
  return B2_OK;
}

}  // extern "C"
