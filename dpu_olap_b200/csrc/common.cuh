// common.cuh — shared device helpers and the context object of libb200olap.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/b200olap.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200olap is written for sm_100a (B200) only"
#endif

// ---- context -----------------------------------------------------------------------------
struct b2_pending;  // device-resident result of the last *_host call (api_host.cu)

struct b2_ctx {
  int device = 0;
  int sm_count = 148;
  int64_t launches = 0;
  std::string last_error;
  // small persistent device scratch: a ring of per-launch slots (partials | ticket) of the sum kernels
  void* d_small = nullptr;
  size_t small_bytes = 0;
  uint32_t sum_slot = 0;  // next slot of that ring
  // kernel-selection knobs (b2_ctx_set_tunable, enum b2_tunable): every setting computes the same result
  int tune[12] = {8, 1, 0, 6, 3, 2048, 0, 0, 0, 0, 0, 0};
  std::vector<char> site_done;  // per call site: function attributes configured for this ctx's device
  std::vector<int> site_value;
  // growable device workspace used by the *_host layer
  void* d_ws = nullptr;
  size_t ws_bytes = 0;
  // host layer
  cudaStream_t s_compute = nullptr;
  cudaStream_t s_copy_in = nullptr;
  cudaStream_t s_copy_out = nullptr;
  void* h_pinned = nullptr;  // pinned staging ring
  size_t pinned_bytes = 0;
  // b2_ctx_set_inputs_pinned: host inputs are page-locked and device-accessible, so a group of
  // batches is uploaded by ONE gather kernel reading host memory instead of one DMA per batch
  bool inputs_pinned = false;
  void* h_gather = nullptr;   // pinned table of gather entries, bump-allocated per *_host call
  size_t gather_bytes = 0, gather_used = 0;
  b2_pending* pending = nullptr;
  // grow-only device buffers of the streaming host entry points (a cudaMalloc + cudaFree of a few
  // GiB per call costs more than the kernels they feed)
  static constexpr int kCacheSlots = 8;
  void* cache_ptr[kCacheSlots] = {};
  size_t cache_bytes[kCacheSlots] = {};
  std::vector<std::pair<void*, size_t>> pool_free;   // idle blocks of the recycling allocator
  std::vector<std::pair<void*, size_t>> pool_live;   // blocks handed out
  // join phase trace (b2_join_trace / b2_join_last_phases): CUDA events at the phase boundaries of the
  // joins launched through this ctx; events are created once and reused
  bool trace_join = false;
  std::vector<cudaEvent_t> trace_events;
  size_t trace_used = 0;
  struct Mark { int phase; size_t ev0, ev1; };
  std::vector<Mark> trace_marks;
};
// Phases of the join trace. Scoped marks: the constructor records the start event on `s`, the
// destructor the end event; nothing happens unless the ctx traces (no events, no host cost).
enum b2_trace_phase { B2_PHASE_PART_BUILD = 0, B2_PHASE_PART_PROBE = 1, B2_PHASE_PROBE = 2, B2_PHASE_TAKE = 3,
                      B2_PHASE_COUNT = 4 };
void b2_trace_reset(b2_ctx* ctx);
struct b2_trace_scope {
  b2_ctx* ctx;
  cudaStream_t s;
  size_t mark = ~(size_t)0;
  b2_trace_scope(b2_ctx* ctx, int phase, cudaStream_t s);
  ~b2_trace_scope();
  b2_trace_scope(const b2_trace_scope&) = delete;
  b2_trace_scope& operator=(const b2_trace_scope&) = delete;
};
// Recycling device allocator of the *_host entry points: b2_dev_free keeps the block for the next
// b2_dev_alloc of a similar size (cudaMalloc / cudaFree of GiB-sized buffers per call cost
// milliseconds and occasionally much more); everything is released in b2_ctx_destroy.
int b2_dev_alloc(b2_ctx* ctx, void** out, size_t bytes);
void b2_dev_free(b2_ctx* ctx, void* p);
// Returns a device buffer of at least `bytes` that stays owned by the ctx (slot 0..kCacheSlots-1).
int b2_ctx_cached(b2_ctx* ctx, int slot, size_t bytes, void** out);

int b2_set_error(b2_ctx* ctx, int status, const char* what, const char* detail);

#define B2_CUDA_OK(ctx, expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      return b2_set_error((ctx), _e == cudaErrorMemoryAllocation ? B2_ERR_OOM : B2_ERR_CUDA, \
                          #expr, cudaGetErrorString(_e));                            \
    }                                                                                \
  } while (0)

#define B2_RETURN_NOT_OK(expr) \
  do {                         \
    int _s = (expr);           \
    if (_s != B2_OK) return _s; \
  } while (0)

#define B2_REQUIRE(ctx, cond, msg)                                        \
  do {                                                                    \
    if (!(cond)) return b2_set_error((ctx), B2_ERR_INVALID, #cond, (msg)); \
  } while (0)

#define B2_LAUNCH_CHECK(ctx, name)                                             \
  do {                                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess)                                                     \
      return b2_set_error((ctx), B2_ERR_CUDA, name, cudaGetErrorString(_e));   \
    (ctx)->launches++;                                                         \
  } while (0)

// Makes ctx->device current for the duration of an entry point and restores the caller's device on
// the way out: kernel launches, function attributes and occupancy queries all act on the CURRENT
// device, and a process may hold contexts on several GPUs (b2_set) or share the thread with torch.
struct b2_device_scope {
  int prev = -1;
  bool switched = false;
  explicit b2_device_scope(const b2_ctx* ctx) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
      switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~b2_device_scope() {
    if (switched) cudaSetDevice(prev);
  }
  b2_device_scope(const b2_device_scope&) = delete;
  b2_device_scope& operator=(const b2_device_scope&) = delete;
};

static inline size_t b2_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Function attributes (opt-in dynamic shared memory) and occupancy are per DEVICE and must be set
// before the first launch there. Every call site that needs them owns a "site" id; the ctx remembers
// which sites it has configured (and one integer per site, e.g. the grid size derived from the
// occupancy query). A ctx is bound to one device and driven by one thread, so this needs no locking
// and is right for processes that hold contexts on several GPUs.
int b2_new_site();  // ctx.cu: process-wide atomic counter
static inline bool b2_first_use_on_device(b2_ctx* ctx, int site) {
  if ((int)ctx->site_done.size() <= site) {
    ctx->site_done.resize((size_t)site + 1, 0);
    ctx->site_value.resize((size_t)site + 1, 0);
  }
  if (ctx->site_done[(size_t)site]) return false;
  ctx->site_done[(size_t)site] = 1;
  return true;
}

// ---- device helpers ----------------------------------------------------------------------
#ifdef __CUDACC__

// Streaming 128-bit load: read-only path, do not allocate in L1 (each byte is touched once).
__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_v2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
  uint64_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// Streaming stores: write-once data, keep it out of L1.
__device__ __forceinline__ void st_stream_v4(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_v2(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Ask the copy engine to bring [p, p + bytes) into L2 (UBLKPF in SASS): no registers, no completion to
// wait for. The range is shrunk to 16-byte alignment at both ends (the edges come with the demand loads).
__device__ __forceinline__ void l2_prefetch(const void* p, int64_t bytes) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uintptr_t a16 = (a + 15) & ~(uintptr_t)15;
  bytes -= (int64_t)(a16 - a);
  bytes &= ~(int64_t)15;
  if (bytes > 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a16), "r"((uint32_t)bytes) : "memory");
}
// Single-word (status|value) tile descriptors of the decoupled look-back scan: a relaxed
// gpu-scope 64-bit access is atomic, so no fence is needed between status and value.
__device__ __forceinline__ uint64_t ld_relaxed_gpu_u64(const uint64_t* p) {
  uint64_t r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Thomas Wang's 32-bit integer hash exactly as the reference uses it for partitioning and for
// its hash table (dpu/shared/kernels/partition.c:20-28, dpu/shared/hashtable/hashtable.c:29-37).
// Written with multiplications: key + ~(key << s) = key * (1 - 2^s) - 1 and key + (key << 3) =
// key * 9 are one IMAD each, which brings the hash from 13 to 9 instructions per row in the
// partitioning kernels (the compiler keeps shift / not / add apart). Same function, bit for bit
// (tests/test_oracle.py pins it to the reference's values).
__host__ __device__ __forceinline__ uint32_t wang_hash_u32(uint32_t key) {
  key = key * 0xFFFF8001u + 0xFFFFFFFFu;  // key += ~(key << 15)
  key ^= (key >> 10);
  key = key * 9u;                         // key += (key << 3)
  key ^= (key >> 6);
  key = key * 0xFFFFF801u + 0xFFFFFFFFu;  // key += ~(key << 11)
  key ^= (key >> 16);
  return key;
}

__device__ __forceinline__ uint64_t warp_reduce_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_reduce_sum_u32(uint32_t v) {
  return __reduce_add_sync(0xffffffffu, v);
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

#endif  // __CUDACC__
