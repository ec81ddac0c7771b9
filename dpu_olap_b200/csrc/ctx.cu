// ctx.cu — context lifetime, status strings. Replaces dpu::DpuSet::allocate/load/~DpuSet
// (reference host/dpuext/dpuext.hpp:669-739): a ctx is bound to one B200 and owns its streams,
// scratch and any pending device-resident result.
#include "common.cuh"

void b2_pending_free(b2_ctx* ctx);  // api_host.cu

#include <atomic>
int b2_new_site() {
  static std::atomic<int> next{0};
  return next.fetch_add(1);
}

int b2_set_error(b2_ctx* ctx, int status, const char* what, const char* detail) {
  if (ctx) {
    ctx->last_error = std::string(b2_strerror(status)) + ": " + (what ? what : "") +
                      (detail ? std::string(" — ") + detail : std::string());
  }
  return status;
}

int b2_dev_alloc(b2_ctx* ctx, void** out, size_t bytes) {
  *out = nullptr;
  if (bytes == 0) bytes = 256;
  // best fit among the idle blocks, but never one more than twice the request
  int best = -1;
  for (size_t i = 0; i < ctx->pool_free.size(); ++i) {
    const size_t cap = ctx->pool_free[i].second;
    if (cap >= bytes && cap <= 2 * bytes + (1 << 20) && (best < 0 || cap < ctx->pool_free[(size_t)best].second))
      best = (int)i;
  }
  if (best >= 0) {
    *out = ctx->pool_free[(size_t)best].first;
    ctx->pool_live.push_back(ctx->pool_free[(size_t)best]);
    ctx->pool_free.erase(ctx->pool_free.begin() + best);
    return B2_OK;
  }
  const size_t want = b2_align_up(bytes, (size_t)1 << 20);
  cudaError_t e = cudaMalloc(out, want);
  if (e != cudaSuccess && !ctx->pool_free.empty()) {  // give the idle blocks back and retry
    cudaGetLastError();
    for (auto& blk : ctx->pool_free) cudaFree(blk.first);
    ctx->pool_free.clear();
    e = cudaMalloc(out, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return b2_set_error(ctx, e == cudaErrorMemoryAllocation ? B2_ERR_OOM : B2_ERR_CUDA, "cudaMalloc",
                        cudaGetErrorString(e));
  }
  ctx->pool_live.emplace_back(*out, want);
  return B2_OK;
}

void b2_dev_free(b2_ctx* ctx, void* p) {
  if (!p) return;
  for (size_t i = 0; i < ctx->pool_live.size(); ++i)
    if (ctx->pool_live[i].first == p) {
      ctx->pool_free.push_back(ctx->pool_live[i]);
      ctx->pool_live.erase(ctx->pool_live.begin() + (long)i);
      return;
    }
  cudaFree(p);  // not from the pool
}

int b2_ctx_cached(b2_ctx* ctx, int slot, size_t bytes, void** out) {
  *out = nullptr;
  if (slot < 0 || slot >= b2_ctx::kCacheSlots) return b2_set_error(ctx, B2_ERR_INVALID, "cache slot", nullptr);
  if (ctx->cache_bytes[slot] < bytes) {
    if (ctx->cache_ptr[slot]) {
      cudaDeviceSynchronize();
      cudaFree(ctx->cache_ptr[slot]);
      ctx->cache_ptr[slot] = nullptr;
      ctx->cache_bytes[slot] = 0;
    }
    const size_t want = b2_align_up(bytes, (size_t)1 << 20);
    B2_CUDA_OK(ctx, cudaMalloc(&ctx->cache_ptr[slot], want));
    ctx->cache_bytes[slot] = want;
  }
  *out = ctx->cache_ptr[slot];
  return B2_OK;
}

// ---- join phase trace ----------------------------------------------------------------------------------
void b2_trace_reset(b2_ctx* ctx) {
  ctx->trace_marks.clear();
  ctx->trace_used = 0;
}
static bool trace_event(b2_ctx* ctx, size_t* idx) {
  if (ctx->trace_used == ctx->trace_events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return false;
    ctx->trace_events.push_back(e);
  }
  *idx = ctx->trace_used++;
  return true;
}
b2_trace_scope::b2_trace_scope(b2_ctx* c, int phase, cudaStream_t st) : ctx(c), s(st) {
  if (!ctx || !ctx->trace_join) return;
  size_t e0, e1;
  if (!trace_event(ctx, &e0) || !trace_event(ctx, &e1)) return;
  if (cudaEventRecord(ctx->trace_events[e0], s) != cudaSuccess) return;
  mark = ctx->trace_marks.size();
  ctx->trace_marks.push_back({phase, e0, e1});
}
b2_trace_scope::~b2_trace_scope() {
  if (mark == ~(size_t)0) return;
  if (cudaEventRecord(ctx->trace_events[ctx->trace_marks[mark].ev1], s) != cudaSuccess)
    ctx->trace_marks[mark].phase = -1;  // never ended: skipped by b2_join_last_phases
}

extern "C" {

int b2_join_trace(b2_ctx* ctx, int on) {
  if (!ctx) return B2_ERR_INVALID;
  ctx->trace_join = on != 0;
  b2_trace_reset(ctx);  // whatever an earlier join left behind is not "the last join" of the new trace
  return B2_OK;
}

int b2_join_last_phases(b2_ctx* ctx, b2_join_phases* out) {
  if (!ctx || !out) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  *out = b2_join_phases{};
  double ms[B2_PHASE_COUNT] = {};
  for (const b2_ctx::Mark& m : ctx->trace_marks) {
    if (m.phase < 0 || m.phase >= B2_PHASE_COUNT) continue;
    float t = 0.f;
    B2_CUDA_OK(ctx, cudaEventSynchronize(ctx->trace_events[m.ev1]));
    B2_CUDA_OK(ctx, cudaEventElapsedTime(&t, ctx->trace_events[m.ev0], ctx->trace_events[m.ev1]));
    ms[m.phase] += (double)t;
    out->intervals++;
  }
  out->partition_build_ms = ms[B2_PHASE_PART_BUILD];
  out->partition_probe_ms = ms[B2_PHASE_PART_PROBE];
  out->probe_ms = ms[B2_PHASE_PROBE];
  out->take_ms = ms[B2_PHASE_TAKE];
  return B2_OK;
}

int b2_version(void) { return B2_VERSION; }

const char* b2_strerror(int status) {
  switch (status) {
    case B2_OK: return "ok";
    case B2_ERR_INVALID: return "invalid argument";
    case B2_ERR_CUDA: return "CUDA error";
    case B2_ERR_OOM: return "out of memory";
    case B2_ERR_UNSUPPORTED: return "unsupported";
    case B2_ERR_WORKSPACE: return "workspace too small";
    case B2_ERR_OVERFLOW: return "output capacity exceeded";
    default: return "unknown status";
  }
}

int b2_device_count(int* count) {
  if (!count) return B2_ERR_INVALID;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return B2_ERR_CUDA;
  }
  *count = n;
  return B2_OK;
}

int b2_ctx_create(int device, b2_ctx** out) {
  if (!out) return B2_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return B2_ERR_CUDA;
  if (device < 0 || device >= n) return B2_ERR_INVALID;
  int prev = -1;
  cudaGetDevice(&prev);
  struct Restore {  // the caller's current device is left as it was (torch shares the thread)
    int prev, device;
    ~Restore() { if (prev >= 0 && prev != device) cudaSetDevice(prev); }
  } restore{prev, device};
  if (cudaSetDevice(device) != cudaSuccess) return B2_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B2_ERR_CUDA;
  if (prop.major < 10) {
    fprintf(stderr, "b200olap: device %d is sm_%d%d; this library is built for sm_100a only\n",
            device, prop.major, prop.minor);
    return B2_ERR_UNSUPPORTED;
  }
  b2_ctx* ctx = new b2_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->small_bytes = 128 * 1024;  // ring of per-launch scratch slots of the sum kernels (sum.cu)
  if (cudaMalloc(&ctx->d_small, ctx->small_bytes) != cudaSuccess ||
      cudaMemset(ctx->d_small, 0, ctx->small_bytes) != cudaSuccess) {
    delete ctx;
    return B2_ERR_OOM;
  }
  *out = ctx;
  return B2_OK;
}

int b2_ctx_destroy(b2_ctx* ctx) {
  if (!ctx) return B2_OK;
  b2_device_scope dev_scope(ctx);
  b2_pending_free(ctx);
  if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if (ctx->s_copy_in) cudaStreamDestroy(ctx->s_copy_in);
  if (ctx->s_copy_out) cudaStreamDestroy(ctx->s_copy_out);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
  if (ctx->d_ws) cudaFree(ctx->d_ws);
  for (void* p : ctx->cache_ptr)
    if (p) cudaFree(p);
  for (auto& blk : ctx->pool_free) cudaFree(blk.first);
  for (auto& blk : ctx->pool_live) cudaFree(blk.first);
  if (ctx->d_small) cudaFree(ctx->d_small);
  for (cudaEvent_t e : ctx->trace_events) cudaEventDestroy(e);
  delete ctx;
  return B2_OK;
}

int b2_ctx_set_inputs_pinned(b2_ctx* ctx, int on) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  ctx->inputs_pinned = on != 0;
  return B2_OK;
}

int b2_host_alloc_pinned(size_t bytes, void** out) {
  if (!out) return B2_ERR_INVALID;
  *out = nullptr;
  if (bytes == 0) bytes = 64;
  const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
  return e == cudaSuccess ? B2_OK : (e == cudaErrorMemoryAllocation ? B2_ERR_OOM : B2_ERR_CUDA);
}
int b2_host_free_pinned(void* p) {
  if (!p) return B2_OK;
  return cudaFreeHost(p) == cudaSuccess ? B2_OK : B2_ERR_CUDA;
}
int b2_host_register(const void* p, size_t bytes) {
  if (!p || bytes == 0) return B2_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable | cudaHostRegisterMapped | cudaHostRegisterReadOnly);
  if (e == cudaSuccess) return B2_OK;
  cudaGetLastError();  // clear the sticky-less error state
  // read-only registration needs driver support; fall back to a plain registration
  const cudaError_t e2 = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e2 == cudaSuccess || e2 == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return B2_OK;
  }
  cudaGetLastError();
  return B2_ERR_CUDA;
}
int b2_host_unregister(const void* p) {
  if (!p) return B2_OK;
  const cudaError_t e = cudaHostUnregister(const_cast<void*>(p));
  cudaGetLastError();
  return (e == cudaSuccess || e == cudaErrorHostMemoryNotRegistered) ? B2_OK : B2_ERR_CUDA;
}

int b2_ctx_set_tunable(b2_ctx* ctx, int which, int value) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  switch (which) {
    case B2_TUNE_SCATTER_SECTORS_MIN_BITS: B2_REQUIRE(ctx, value >= 0, "0 = always, > 10 = never"); break;
    case B2_TUNE_SCATTER_PREFETCH: value = value != 0; break;
    case B2_TUNE_SCATTER_SHAPE: B2_REQUIRE(ctx, (value >= 0 && value <= 3) || value == 8, "shape 0..3 or 8"); break;
    case B2_TUNE_FILTER_VARIANT: B2_REQUIRE(ctx, value >= 0 && value <= 7, "variant 0..7"); break;
    case B2_TUNE_SCATTER_SECTOR_TILE: B2_REQUIRE(ctx, value >= 0 && value <= 3, "0 .. 3"); break;
    case B2_TUNE_PEER_SCATTER_KERNEL: B2_REQUIRE(ctx, value == 0 || value == 1, "0 = lines, 1 = bulk sectors"); break;
    case B2_TUNE_PEER_SCATTER_CTAS: B2_REQUIRE(ctx, value >= 0, "0 = one CTA per work unit"); break;
    case B2_TUNE_JOIN_DIRECT_MIN_ROWS: B2_REQUIRE(ctx, value >= 0, "0 = off, n = minimum build rows per partition"); break;
    case B2_TUNE_FILTER64_KERNEL: B2_REQUIRE(ctx, value == 0 || value == 1, "0 = single pass, 1 = two passes"); break;
    default: return b2_set_error(ctx, B2_ERR_INVALID, "b2_ctx_set_tunable", "unknown tunable");
  }
  ctx->tune[which] = value;
  return B2_OK;
}
int b2_ctx_get_tunable(const b2_ctx* ctx, int which, int* value) {
  if (!ctx || !value || which < 0 || which > B2_TUNE_FILTER64_KERNEL) return B2_ERR_INVALID;
  *value = ctx->tune[which];
  return B2_OK;
}

const char* b2_last_error(const b2_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
int64_t b2_launch_count(const b2_ctx* ctx) { return ctx ? ctx->launches : 0; }
int b2_ctx_device(const b2_ctx* ctx) { return ctx ? ctx->device : -1; }
int b2_ctx_sm_count(const b2_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
