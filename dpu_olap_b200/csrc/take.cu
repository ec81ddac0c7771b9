// take.cu — batch-local gather out[b][j] = values[b][indices[b][j]] over uint32 columns.
//
// Replaces the reference's DPU take program (dpu/shared/kernels/take.c:12-47): each tasklet reads
// a block of selection indices into WRAM and fetches every value with an 8-byte random MRAM DMA
// (take.c:37), without bounds checking (:36). The Acero side is cp::Take(values, indices,
// NoBoundsCheck) per batch (host/take/take_native.cc:24-31).
//
// B200 design: indices and output are streamed with 128-bit accesses (4 indices per load, 4
// loads = 16 independent gathers in flight per thread); the gathers go through the read-only
// path. A batch's value window (16 MiB at the benchmark shape) is far smaller than the 126 MB L2
// and consecutive CTAs work on the same batch, so every touched 32 B sector is fetched from HBM
// once. HBM traffic is therefore bounded by sectors touched, not by 4 B per gather.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;  // uint4 index vectors per thread

__device__ __forceinline__ uint32_t gather(const uint32_t* __restrict__ p) { return __ldg(p); }

// Uniform batches, idx_len % 4 == 0, 16 B aligned index/out pointers.
__global__ void __launch_bounds__(kThreads)
take_u32_vec_kernel(const uint32_t* __restrict__ values, int64_t values_len,
                    const uint4* __restrict__ indices, int64_t idx_len, int64_t nvec,
                    uint4* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  uint4 ix[kUnroll];
  uint4 r[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) ix[u] = ld_stream_v4(indices + i);
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) {
      const int64_t b = (i << 2) / idx_len;
      const uint32_t* __restrict__ v = values + b * values_len;
      r[u].x = gather(v + ix[u].x);
      r[u].y = gather(v + ix[u].y);
      r[u].z = gather(v + ix[u].z);
      r[u].w = gather(v + ix[u].w);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) st_stream_v4(out + i, r[u]);
  }
}

// General path: any lengths/alignment; optional ragged offset tables.
__global__ void __launch_bounds__(kThreads)
take_u32_scalar_kernel(const uint32_t* __restrict__ values, int64_t values_len,
                       const int64_t* __restrict__ values_off, const uint32_t* __restrict__ indices,
                       int64_t idx_len, const int64_t* __restrict__ idx_off, int64_t nbatches,
                       int64_t i0, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = i0 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kThreads) {
    int64_t vbase;
    if (idx_off) {
      int64_t lo = 0, hi = nbatches;  // last b with idx_off[b] <= i
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (idx_off[mid] <= i) lo = mid; else hi = mid;
      }
      vbase = values_off[lo];
    } else {
      vbase = (i / idx_len) * values_len;
    }
    out[i] = gather(values + vbase + indices[i]);
  }
}

// ---- 64-bit values (uint64 / int64 / float64 as raw words), 32-bit indices ----------------------
// Four indices per 128-bit load, four 8-byte gathers, two 128-bit stores per index vector.
__global__ void __launch_bounds__(kThreads)
take_u64_vec_kernel(const unsigned long long* __restrict__ values, int64_t values_len,
                    const uint4* __restrict__ indices, int64_t idx_len, int64_t nvec,
                    uint4* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  uint4 ix[kUnroll];
  unsigned long long r[kUnroll][4];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) ix[u] = ld_stream_v4(indices + i);
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) {
      const int64_t b = (i << 2) / idx_len;
      const unsigned long long* __restrict__ v = values + b * values_len;
      r[u][0] = __ldg(v + ix[u].x);
      r[u][1] = __ldg(v + ix[u].y);
      r[u][2] = __ldg(v + ix[u].z);
      r[u][3] = __ldg(v + ix[u].w);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + u * kThreads;
    if (i < nvec) {
      st_stream_v4(out + 2 * i, make_uint4((uint32_t)r[u][0], (uint32_t)(r[u][0] >> 32), (uint32_t)r[u][1],
                                           (uint32_t)(r[u][1] >> 32)));
      st_stream_v4(out + 2 * i + 1, make_uint4((uint32_t)r[u][2], (uint32_t)(r[u][2] >> 32), (uint32_t)r[u][3],
                                               (uint32_t)(r[u][3] >> 32)));
    }
  }
}

// Any lengths / alignment (uniform batches).
__global__ void __launch_bounds__(kThreads)
take_u64_scalar_kernel(const unsigned long long* __restrict__ values, int64_t values_len,
                       const uint32_t* __restrict__ indices, int64_t idx_len, int64_t n,
                       unsigned long long* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    out[i] = __ldg(values + (i / idx_len) * values_len + indices[i]);
}

}  // namespace

extern "C" {

int b2_take_64_dev(b2_ctx* ctx, const void* d_values_, int64_t values_len, const uint32_t* d_indices,
                   int64_t idx_len, int64_t nbatches, void* d_out_, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, values_len >= 0 && idx_len >= 0 && nbatches >= 0, "negative size");
  const int64_t n = nbatches * idx_len;
  if (n == 0) return B2_OK;
  const unsigned long long* d_values = static_cast<const unsigned long long*>(d_values_);
  unsigned long long* d_out = static_cast<unsigned long long*>(d_out_);
  B2_REQUIRE(ctx, d_values && d_indices && d_out, "null column pointer");
  B2_REQUIRE(ctx, values_len > 0, "indices into an empty values batch");
  B2_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_values) | reinterpret_cast<uintptr_t>(d_out)) & 7) == 0,
             "64-bit columns must be 8-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (idx_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_indices) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
  if (vec) {
    const int64_t nvec = n >> 2;
    const int64_t per_cta = (int64_t)kThreads * kUnroll;
    const int64_t grid = (nvec + per_cta - 1) / per_cta;
    B2_REQUIRE(ctx, grid < (1ll << 31), "too many indices for one launch");
    take_u64_vec_kernel<<<(unsigned)grid, kThreads, 0, s>>>(d_values, values_len,
                                                            reinterpret_cast<const uint4*>(d_indices), idx_len,
                                                            nvec, reinterpret_cast<uint4*>(d_out));
    B2_LAUNCH_CHECK(ctx, "take_u64_vec_kernel");
  } else {
    int64_t grid = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    if (grid > cap) grid = cap;
    take_u64_scalar_kernel<<<(unsigned)grid, kThreads, 0, s>>>(d_values, values_len, d_indices, idx_len, n, d_out);
    B2_LAUNCH_CHECK(ctx, "take_u64_scalar_kernel");
  }
  return B2_OK;
}

int b2_take_u32_dev(b2_ctx* ctx, const uint32_t* d_values, int64_t values_len,
                    const uint32_t* d_indices, int64_t idx_len, int64_t nbatches, uint32_t* d_out,
                    void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, values_len >= 0 && idx_len >= 0 && nbatches >= 0, "negative size");
  const int64_t n = nbatches * idx_len;
  if (n == 0) return B2_OK;
  B2_REQUIRE(ctx, d_values && d_indices && d_out, "null column pointer");
  B2_REQUIRE(ctx, values_len > 0, "indices into an empty values batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (idx_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_indices) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
  if (vec) {
    const int64_t nvec = n >> 2;
    const int64_t per_cta = (int64_t)kThreads * kUnroll;
    const int64_t grid = (nvec + per_cta - 1) / per_cta;
    B2_REQUIRE(ctx, grid < (1ll << 31), "too many indices for one launch");
    take_u32_vec_kernel<<<(unsigned)grid, kThreads, 0, s>>>(
        d_values, values_len, reinterpret_cast<const uint4*>(d_indices), idx_len, nvec,
        reinterpret_cast<uint4*>(d_out));
    B2_LAUNCH_CHECK(ctx, "take_u32_vec_kernel");
  } else {
    int64_t grid = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    if (grid > cap) grid = cap;
    take_u32_scalar_kernel<<<(unsigned)grid, kThreads, 0, s>>>(d_values, values_len, nullptr,
                                                               d_indices, idx_len, nullptr,
                                                               nbatches, 0, n, d_out);
    B2_LAUNCH_CHECK(ctx, "take_u32_scalar_kernel");
  }
  return B2_OK;
}

int b2_take_u32_ragged_dev(b2_ctx* ctx, const uint32_t* d_values, const int64_t* d_values_off,
                           const uint32_t* d_indices, const int64_t* d_idx_off, int64_t nbatches,
                           int64_t idx_begin, int64_t idx_end, uint32_t* d_out, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches >= 0 && idx_begin >= 0 && idx_end >= idx_begin, "bad range");
  const int64_t n_idx_total = idx_end - idx_begin;
  if (n_idx_total == 0) return B2_OK;
  B2_REQUIRE(ctx, d_values && d_indices && d_out && d_values_off && d_idx_off, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t grid = (n_idx_total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)ctx->sm_count * 32;
  if (grid > cap) grid = cap;
  take_u32_scalar_kernel<<<(unsigned)grid, kThreads, 0, s>>>(d_values, 0, d_values_off, d_indices,
                                                             0, d_idx_off, nbatches, idx_begin,
                                                             idx_end, d_out);
  B2_LAUNCH_CHECK(ctx, "take_u32_scalar_kernel");
  return B2_OK;
}

}  // extern "C"
