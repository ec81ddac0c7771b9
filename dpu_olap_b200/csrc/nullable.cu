// nullable.cu — aggregate and take over columns WITH validity bitmaps.
//
// The reference's DPU path has neither: its kernels see raw uint32 buffers only (it hands Arrow a
// nullptr bitmap, host/filter/filter_dpu.cc:91; `enum AggregatorType` holds only AggrSum,
// shared/umq/kernels.h:22-25). What is restated here is the behaviour of the reference's ORACLE,
// Arrow's compute kernels, which its Native classes call (aggr_native.cc:68-73, take_native.cc:27):
//   * aggregates skip null rows: sum -> uint64, count = valid rows, min / max over valid rows
//     (all four in one pass: the kernel is a 4 B/row read stream either way);
//   * take: out[j] is null when the INDEX j is null or the VALUE it points at is null.
// Bitmaps are Arrow's: bit i of the buffer = row i, least significant bit first; here one bitmap
// spans the packed column (row i of the packed uint32[]), 4-byte aligned and padded to 4 bytes.
// The nullable filter is a variant of the filter kernel itself (filter.cu).
#include "common.cuh"

namespace {

constexpr int kAggThreads = 512;
constexpr int kAggUnroll = 8;  // 128-bit loads in flight per thread, as the sum kernel

// kSigned: the column is int32 — sum into int64 (sign-extended), min / max in signed order; the
// results travel as bit patterns in the same b2_aggr_u32 fields.
template <bool kSigned>
struct AggAcc {
  uint64_t sum = 0;
  uint32_t cnt = 0;  // flushed into the 64-bit total by the caller's reduction
  uint32_t mn = kSigned ? 0x7fffffffu : 0xffffffffu;
  uint32_t mx = kSigned ? 0x80000000u : 0u;
};

template <bool kSigned>
__device__ __forceinline__ void agg_row(AggAcc<kSigned>& a, uint32_t v, bool valid) {
  if (valid) {
    a.cnt += 1;
    if (kSigned) {
      a.sum += (uint64_t)(int64_t)(int32_t)v;
      a.mn = (uint32_t)min((int32_t)a.mn, (int32_t)v);
      a.mx = (uint32_t)max((int32_t)a.mx, (int32_t)v);
    } else {
      a.sum += v;
      a.mn = min(a.mn, v);
      a.mx = max(a.mx, v);
    }
  }
}

// One pass: 128-bit value loads, one 32-bit bitmap word per 8 threads (valid == nullptr: no nulls).
// The first `head` rows (until 16 B alignment AND a nibble boundary of the bitmap) and the tail are
// done row by row by CTA 0.
template <bool kHasValid, bool kSigned>
__global__ void __launch_bounds__(kAggThreads, 2)
aggr_u32_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ valid_, int64_t n,
                int64_t head, b2_aggr_u32* __restrict__ out) {
  const uint32_t* __restrict__ valid = kHasValid ? valid_ : nullptr;
  const int64_t nvec = (n - head) >> 2;
  const int64_t tail_start = head + (nvec << 2);
  const uint4* __restrict__ vin = reinterpret_cast<const uint4*>(in + head);
  AggAcc<kSigned> acc;
  uint64_t cnt = 0;
  const int64_t chunk = (int64_t)kAggThreads * kAggUnroll;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    uint4 v[kAggUnroll];
    uint32_t nib[kAggUnroll];
    if (base + chunk <= nvec) {
      // full chunk, straight-line: 8 value loads and 8 bitmap words in flight before the first use
      // (with a branch per load the compiler consumed every bitmap word right after its load: eight
      // serial L2 round trips per iteration, 3.9 TB/s)
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) v[u] = ld_stream_v4(vin + base + u * kAggThreads + threadIdx.x);
      if (kHasValid) {
        // head == 0 with a bitmap: vector i covers rows 4i..4i+3 = nibble (i & 7) of word i >> 3
        const uint32_t* __restrict__ wp = valid + ((base + threadIdx.x) >> 3);
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) nib[u] = __ldg(wp + u * (kAggThreads / 8));
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) nib[u] = (nib[u] >> ((threadIdx.x & 7) * 4)) & 0xfu;
      } else {
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) nib[u] = 0xfu;
      }
    } else {
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) {
        const int64_t i = base + u * kAggThreads + threadIdx.x;
        nib[u] = 0;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i < nvec) {
          v[u] = ld_stream_v4(vin + i);
          const int64_t r = head + (i << 2);
          nib[u] = kHasValid ? (__ldg(valid + (r >> 5)) >> (r & 31)) & 0xfu : 0xfu;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kAggUnroll; ++u) {
      agg_row(acc, v[u].x, nib[u] & 1u);
      agg_row(acc, v[u].y, nib[u] & 2u);
      agg_row(acc, v[u].z, nib[u] & 4u);
      agg_row(acc, v[u].w, nib[u] & 8u);
    }
    cnt += acc.cnt;
    acc.cnt = 0;
  }
  // rows outside the vector body: all of them when the column cannot be vectorised (head == n,
  // shared by the whole grid), else the few head / tail rows (CTA 0)
  auto scalar_rows = [&](int64_t r0, int64_t r1, int64_t first, int64_t step) {
    for (int64_t i = r0 + first; i < r1; i += step)
      agg_row(acc, in[i], valid ? (__ldg(valid + (i >> 5)) >> (i & 31)) & 1u : 1u);
  };
  if (nvec == 0) {
    scalar_rows(0, n, (int64_t)blockIdx.x * kAggThreads + threadIdx.x, (int64_t)gridDim.x * kAggThreads);
  } else if (blockIdx.x == 0) {
    scalar_rows(0, head, threadIdx.x, kAggThreads);
    scalar_rows(tail_start, n, threadIdx.x, kAggThreads);
  }
  cnt += acc.cnt;
  // CTA reduction, then four atomics per CTA on the (pre-initialised) result
  __shared__ uint64_t s_sum[kAggThreads / 32], s_cnt[kAggThreads / 32];
  __shared__ uint32_t s_mn[kAggThreads / 32], s_mx[kAggThreads / 32];
  const uint64_t wsum = warp_reduce_sum_u64(acc.sum), wcnt = warp_reduce_sum_u64(cnt);
  const uint32_t wmn = kSigned ? (uint32_t)__reduce_min_sync(0xffffffffu, (int32_t)acc.mn)
                               : __reduce_min_sync(0xffffffffu, acc.mn);
  const uint32_t wmx = kSigned ? (uint32_t)__reduce_max_sync(0xffffffffu, (int32_t)acc.mx)
                               : __reduce_max_sync(0xffffffffu, acc.mx);
  if (lane_id() == 0) {
    s_sum[threadIdx.x >> 5] = wsum;
    s_cnt[threadIdx.x >> 5] = wcnt;
    s_mn[threadIdx.x >> 5] = wmn;
    s_mx[threadIdx.x >> 5] = wmx;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool in_range = threadIdx.x < kAggThreads / 32;
    const uint64_t a = warp_reduce_sum_u64(in_range ? s_sum[threadIdx.x] : 0);
    const uint64_t c = warp_reduce_sum_u64(in_range ? s_cnt[threadIdx.x] : 0);
    const AggAcc<kSigned> none;  // identity elements of min / max for lanes without a warp
    const uint32_t mn_l = in_range ? s_mn[threadIdx.x] : none.mn, mx_l = in_range ? s_mx[threadIdx.x] : none.mx;
    const uint32_t lo = kSigned ? (uint32_t)__reduce_min_sync(0xffffffffu, (int32_t)mn_l)
                                : __reduce_min_sync(0xffffffffu, mn_l);
    const uint32_t hi = kSigned ? (uint32_t)__reduce_max_sync(0xffffffffu, (int32_t)mx_l)
                                : __reduce_max_sync(0xffffffffu, mx_l);
    if (threadIdx.x == 0 && c > 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&out->sum), (unsigned long long)a);
      atomicAdd(reinterpret_cast<unsigned long long*>(&out->count), (unsigned long long)c);
      if (kSigned) {
        atomicMin(reinterpret_cast<int*>(&out->min), (int)lo);
        atomicMax(reinterpret_cast<int*>(&out->max), (int)hi);
      } else {
        atomicMin(&out->min, lo);
        atomicMax(&out->max, hi);
      }
    }
  }
}

__global__ void aggr_init_kernel(b2_aggr_u32* out, bool is_signed) {
  out->sum = 0;
  out->count = 0;
  out->min = is_signed ? 0x7fffffffu : 0xffffffffu;
  out->max = is_signed ? 0x80000000u : 0u;
}

// ---- 64-bit columns ----------------------------------------------------------------------------
// Same pass over uint64 / int64 values: two values per 128-bit load, two validity bits per vector
// (vector i covers rows 2i, 2i+1 = bits 2*(i & 15) of bitmap word i >> 4). The sum wraps mod 2^64
// in both types (Arrow's sum is unchecked); min / max compare in the type's order.
template <bool kSigned>
struct AggAcc64 {
  uint64_t sum = 0;
  uint32_t cnt = 0;
  uint64_t mn = kSigned ? 0x7fffffffffffffffull : 0xffffffffffffffffull;
  uint64_t mx = kSigned ? 0x8000000000000000ull : 0ull;
};
template <bool kSigned>
__device__ __forceinline__ uint64_t min64(uint64_t a, uint64_t b) {
  return kSigned ? (uint64_t)min((long long)a, (long long)b) : min(a, b);
}
template <bool kSigned>
__device__ __forceinline__ uint64_t max64(uint64_t a, uint64_t b) {
  return kSigned ? (uint64_t)max((long long)a, (long long)b) : max(a, b);
}
template <bool kSigned>
__device__ __forceinline__ void agg_row64(AggAcc64<kSigned>& a, uint64_t v, bool valid) {
  if (valid) {
    a.cnt += 1;
    a.sum += v;
    a.mn = min64<kSigned>(a.mn, v);
    a.mx = max64<kSigned>(a.mx, v);
  }
}
template <bool kSigned>
__device__ __forceinline__ uint64_t warp_min64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min64<kSigned>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <bool kSigned>
__device__ __forceinline__ uint64_t warp_max64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max64<kSigned>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <bool kHasValid, bool kSigned>
__global__ void __launch_bounds__(kAggThreads, 2)
aggr_u64_kernel(const uint64_t* __restrict__ in, const uint32_t* __restrict__ valid_, int64_t n,
                int64_t head, b2_aggr_u64* __restrict__ out) {
  const uint32_t* __restrict__ valid = kHasValid ? valid_ : nullptr;
  const int64_t nvec = (n - head) >> 1;
  const int64_t tail_start = head + (nvec << 1);
  const uint4* __restrict__ vin = reinterpret_cast<const uint4*>(in + head);
  AggAcc64<kSigned> acc;
  uint64_t cnt = 0;
  const int64_t chunk = (int64_t)kAggThreads * kAggUnroll;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    uint4 v[kAggUnroll];
    uint32_t two[kAggUnroll];
    if (base + chunk <= nvec) {
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) v[u] = ld_stream_v4(vin + base + u * kAggThreads + threadIdx.x);
      if (kHasValid) {
        // head == 0 with a bitmap; base is a multiple of 16 vectors
        const uint32_t* __restrict__ wp = valid + ((base + threadIdx.x) >> 4);
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) two[u] = __ldg(wp + u * (kAggThreads / 16));
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) two[u] = (two[u] >> ((threadIdx.x & 15) * 2)) & 3u;
      } else {
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) two[u] = 3u;
      }
    } else {
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) {
        const int64_t i = base + u * kAggThreads + threadIdx.x;
        two[u] = 0;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i < nvec) {
          v[u] = ld_stream_v4(vin + i);
          const int64_t r = head + (i << 1);
          two[u] = kHasValid ? (__ldg(valid + (r >> 5)) >> (r & 31)) & 3u : 3u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kAggUnroll; ++u) {
      agg_row64(acc, (uint64_t)v[u].x | ((uint64_t)v[u].y << 32), two[u] & 1u);
      agg_row64(acc, (uint64_t)v[u].z | ((uint64_t)v[u].w << 32), two[u] & 2u);
    }
    cnt += acc.cnt;
    acc.cnt = 0;
  }
  auto scalar_rows = [&](int64_t r0, int64_t r1, int64_t first, int64_t step) {
    for (int64_t i = r0 + first; i < r1; i += step)
      agg_row64(acc, in[i], valid ? (__ldg(valid + (i >> 5)) >> (i & 31)) & 1u : 1u);
  };
  if (nvec == 0) {
    scalar_rows(0, n, (int64_t)blockIdx.x * kAggThreads + threadIdx.x, (int64_t)gridDim.x * kAggThreads);
  } else if (blockIdx.x == 0) {
    scalar_rows(0, head, threadIdx.x, kAggThreads);
    scalar_rows(tail_start, n, threadIdx.x, kAggThreads);
  }
  cnt += acc.cnt;
  __shared__ uint64_t s_sum[kAggThreads / 32], s_cnt[kAggThreads / 32], s_mn[kAggThreads / 32], s_mx[kAggThreads / 32];
  const uint64_t wsum = warp_reduce_sum_u64(acc.sum), wcnt = warp_reduce_sum_u64(cnt);
  const uint64_t wmn = warp_min64<kSigned>(acc.mn), wmx = warp_max64<kSigned>(acc.mx);
  if (lane_id() == 0) {
    s_sum[threadIdx.x >> 5] = wsum;
    s_cnt[threadIdx.x >> 5] = wcnt;
    s_mn[threadIdx.x >> 5] = wmn;
    s_mx[threadIdx.x >> 5] = wmx;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool in_range = threadIdx.x < kAggThreads / 32;
    const AggAcc64<kSigned> none;  // identity elements of min / max for lanes without a warp
    const uint64_t a = warp_reduce_sum_u64(in_range ? s_sum[threadIdx.x] : 0);
    const uint64_t c = warp_reduce_sum_u64(in_range ? s_cnt[threadIdx.x] : 0);
    const uint64_t lo = warp_min64<kSigned>(in_range ? s_mn[threadIdx.x] : none.mn);
    const uint64_t hi = warp_max64<kSigned>(in_range ? s_mx[threadIdx.x] : none.mx);
    if (threadIdx.x == 0 && c > 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&out->sum), (unsigned long long)a);
      atomicAdd(reinterpret_cast<unsigned long long*>(&out->count), (unsigned long long)c);
      if (kSigned) {
        atomicMin(reinterpret_cast<long long*>(&out->min), (long long)lo);
        atomicMax(reinterpret_cast<long long*>(&out->max), (long long)hi);
      } else {
        atomicMin(reinterpret_cast<unsigned long long*>(&out->min), (unsigned long long)lo);
        atomicMax(reinterpret_cast<unsigned long long*>(&out->max), (unsigned long long)hi);
      }
    }
  }
}

__global__ void aggr64_init_kernel(b2_aggr_u64* out, bool is_signed) {
  out->sum = 0;
  out->count = 0;
  out->min = is_signed ? 0x7fffffffffffffffull : 0xffffffffffffffffull;
  out->max = is_signed ? 0x8000000000000000ull : 0ull;
}

// ---- take ------------------------------------------------------------------------------------
constexpr int kTakeThreads = 256;
constexpr int kTakeWords = 8;  // 32-output words per warp and iteration (independent gathers in flight)

__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ bm, int64_t i) {
  return (__ldg(bm + (i >> 5)) >> (i & 31)) & 1u;
}

// A warp owns whole 32-output words of the result bitmap: lane l handles output 32*w + l, the
// validity word is one ballot. Null slots carry the value 0.
template <bool kValValid, bool kIdxValid>
__global__ void __launch_bounds__(kTakeThreads)
take_u32_nullable_kernel(const uint32_t* __restrict__ values, const uint32_t* __restrict__ values_valid,
                         int64_t values_len, const uint32_t* __restrict__ indices,
                         const uint32_t* __restrict__ indices_valid, int64_t idx_len, int64_t n,
                         uint32_t* __restrict__ out, uint32_t* __restrict__ out_valid) {
  const uint32_t lane = threadIdx.x & 31;
  const int64_t nwords = (n + 31) >> 5;
  const int64_t warp0 = ((int64_t)blockIdx.x * kTakeThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kTakeThreads) >> 5;
  for (int64_t w0 = warp0 * kTakeWords; w0 < nwords; w0 += nwarps * kTakeWords) {
    uint32_t ix[kTakeWords], val[kTakeWords];
    bool ok[kTakeWords];
    // straight-line: all index loads, then all gathers (value and its validity word together), so a
    // thread has 2 x kTakeWords independent random loads in flight
    uint32_t iw[kTakeWords];
#pragma unroll
    for (int u = 0; u < kTakeWords; ++u) {
      const int64_t i = ((w0 + u) << 5) + lane;
      ok[u] = i < n;
      ix[u] = ld_stream_u32(indices + (ok[u] ? i : 0));  // clamped, not branched: the loads stay batched
      iw[u] = kIdxValid ? __ldg(indices_valid + min(w0 + u, nwords - 1)) : 0xffffffffu;
    }
    uint32_t vw[kTakeWords];
    int64_t r[kTakeWords];
    // batch of output i: ONE 64-bit division per thread and iteration, then +32 outputs per word
    const int64_t i0 = (w0 << 5) + lane;
    int64_t vbase = (i0 / idx_len) * values_len;  // first row of the batch in the packed values
    int64_t rem = i0 % idx_len;                   // position inside the batch
#pragma unroll
    for (int u = 0; u < kTakeWords; ++u) {
      ok[u] = ok[u] && ((iw[u] >> lane) & 1u);
      r[u] = ok[u] ? vbase + ix[u] : 0;  // batch-local gather (take_native.cc:27)
      val[u] = __ldg(values + r[u]);
      vw[u] = kValValid ? __ldg(values_valid + (r[u] >> 5)) : 0xffffffffu;
      rem += 32;
      while (rem >= idx_len) {
        rem -= idx_len;
        vbase += values_len;
      }
    }
#pragma unroll
    for (int u = 0; u < kTakeWords; ++u) {
      ok[u] = ok[u] && ((vw[u] >> (r[u] & 31)) & 1u);
      if (!ok[u]) val[u] = 0;
    }
#pragma unroll
    for (int u = 0; u < kTakeWords; ++u) {
      const int64_t i = ((w0 + u) << 5) + lane;
      const uint32_t word = __ballot_sync(0xffffffffu, ok[u]);
      if (i < n) st_stream_u32(out + i, val[u]);
      if (out_valid && lane == 0 && w0 + u < nwords) out_valid[w0 + u] = word;
    }
  }
}

}  // namespace

extern "C" {

int b2_aggr_u32_dev(b2_ctx* ctx, const uint32_t* d_in, const uint8_t* d_valid, int64_t n,
                    b2_aggr_u32* d_out, void* stream) {
  return b2_aggr_32_dev(ctx, d_in, B2_U32, d_valid, n, d_out, stream);
}

int b2_aggr_32_dev(b2_ctx* ctx, const void* d_in_, int dtype, const uint8_t* d_valid, int64_t n,
                   b2_aggr_u32* d_out, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  if (dtype == B2_F32)
    return b2_set_error(ctx, B2_ERR_UNSUPPORTED, "float32 aggregates",
                        "rounding and the sign of zero depend on the summation order, in Arrow too");
  B2_REQUIRE(ctx, dtype == B2_U32 || dtype == B2_I32, "dtype must be B2_U32 or B2_I32");
  const bool is_signed = dtype == B2_I32;
  const uint32_t* d_in = static_cast<const uint32_t*>(d_in_);
  B2_REQUIRE(ctx, n >= 0, "n must be >= 0");
  B2_REQUIRE(ctx, d_out != nullptr, "d_out is null");
  B2_REQUIRE(ctx, n == 0 || d_in != nullptr, "d_in is null");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 3) == 0, "d_in must be 4-byte aligned");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_valid) & 3) == 0, "validity bitmap must be 4-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  aggr_init_kernel<<<1, 1, 0, s>>>(d_out, is_signed);
  B2_LAUNCH_CHECK(ctx, "aggr_init_kernel");
  if (n == 0) return B2_OK;
  // scalar head: rows until the values are 16 B aligned; with a bitmap the vector body must also
  // start on a multiple of 4 rows, which a 16 B-aligned start of a 4 B-aligned array always is
  // when the array itself starts on a 16 B boundary — otherwise fall back to all-scalar.
  const uintptr_t addr = reinterpret_cast<uintptr_t>(d_in);
  int64_t head = (int64_t)(((16 - (addr & 15)) & 15) >> 2);
  if (d_valid && head != 0) head = n;  // misaligned values with a bitmap: row-by-row path
  if (head > n) head = n;
  const int64_t chunk_rows = (int64_t)kAggThreads * kAggUnroll * 4;
  int64_t want = (n - head + chunk_rows - 1) / chunk_rows;
  if (head == n) want = (n + kAggThreads - 1) / kAggThreads;  // row-by-row path: one row per thread and step
  int grid = ctx->sm_count * 2;
  if (want < grid) grid = want > 0 ? (int)want : 1;
  const uint32_t* vb = reinterpret_cast<const uint32_t*>(d_valid);
  if (d_valid && is_signed) aggr_u32_kernel<true, true><<<grid, kAggThreads, 0, s>>>(d_in, vb, n, head, d_out);
  else if (d_valid) aggr_u32_kernel<true, false><<<grid, kAggThreads, 0, s>>>(d_in, vb, n, head, d_out);
  else if (is_signed) aggr_u32_kernel<false, true><<<grid, kAggThreads, 0, s>>>(d_in, nullptr, n, head, d_out);
  else aggr_u32_kernel<false, false><<<grid, kAggThreads, 0, s>>>(d_in, nullptr, n, head, d_out);
  B2_LAUNCH_CHECK(ctx, "aggr_u32_kernel");
  return B2_OK;
}

int b2_aggr_64_dev(b2_ctx* ctx, const void* d_in_, int dtype, const uint8_t* d_valid, int64_t n,
                   b2_aggr_u64* d_out, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, dtype == B2_U64 || dtype == B2_I64, "dtype must be B2_U64 or B2_I64");
  const bool is_signed = dtype == B2_I64;
  const uint64_t* d_in = static_cast<const uint64_t*>(d_in_);
  B2_REQUIRE(ctx, n >= 0, "n must be >= 0");
  B2_REQUIRE(ctx, d_out != nullptr, "d_out is null");
  B2_REQUIRE(ctx, n == 0 || d_in != nullptr, "d_in is null");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 7) == 0, "d_in must be 8-byte aligned");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_valid) & 3) == 0, "validity bitmap must be 4-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  aggr64_init_kernel<<<1, 1, 0, s>>>(d_out, is_signed);
  B2_LAUNCH_CHECK(ctx, "aggr64_init_kernel");
  if (n == 0) return B2_OK;
  // scalar head: one row when the values start 8 bytes off a 16-byte boundary; with a bitmap the
  // vector body must start on an even row, so a misaligned start takes the row-by-row path
  int64_t head = (reinterpret_cast<uintptr_t>(d_in) & 15) ? 1 : 0;
  if (d_valid && head != 0) head = n;
  if (head > n) head = n;
  const int64_t chunk_rows = (int64_t)kAggThreads * kAggUnroll * 2;
  int64_t want = (n - head + chunk_rows - 1) / chunk_rows;
  if (head == n) want = (n + kAggThreads - 1) / kAggThreads;
  int grid = ctx->sm_count * 2;
  if (want < grid) grid = want > 0 ? (int)want : 1;
  const uint32_t* vb = reinterpret_cast<const uint32_t*>(d_valid);
  if (d_valid && is_signed) aggr_u64_kernel<true, true><<<grid, kAggThreads, 0, s>>>(d_in, vb, n, head, d_out);
  else if (d_valid) aggr_u64_kernel<true, false><<<grid, kAggThreads, 0, s>>>(d_in, vb, n, head, d_out);
  else if (is_signed) aggr_u64_kernel<false, true><<<grid, kAggThreads, 0, s>>>(d_in, nullptr, n, head, d_out);
  else aggr_u64_kernel<false, false><<<grid, kAggThreads, 0, s>>>(d_in, nullptr, n, head, d_out);
  B2_LAUNCH_CHECK(ctx, "aggr_u64_kernel");
  return B2_OK;
}

int b2_take_u32_nullable_dev(b2_ctx* ctx, const uint32_t* d_values, const uint8_t* d_values_valid,
                             int64_t values_len, const uint32_t* d_indices, const uint8_t* d_indices_valid,
                             int64_t idx_len, int64_t nbatches, uint32_t* d_out, uint8_t* d_out_valid,
                             void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, values_len >= 0 && idx_len >= 0 && nbatches >= 0, "negative size");
  const int64_t n = nbatches * idx_len;
  if (n == 0) return B2_OK;
  B2_REQUIRE(ctx, d_values && d_indices && d_out, "null column pointer");
  B2_REQUIRE(ctx, values_len > 0, "indices into an empty values batch");
  B2_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_values_valid) | reinterpret_cast<uintptr_t>(d_indices_valid) |
                    reinterpret_cast<uintptr_t>(d_out_valid)) & 3) == 0,
             "validity bitmaps must be 4-byte aligned");
  B2_REQUIRE(ctx, d_out_valid || (!d_values_valid && !d_indices_valid),
             "nullable inputs need an output validity bitmap");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nwords = (n + 31) >> 5;
  const int64_t warps_per_cta = kTakeThreads / 32;
  int64_t grid = (nwords + warps_per_cta * kTakeWords - 1) / (warps_per_cta * kTakeWords);
  const int64_t cap = (int64_t)ctx->sm_count * 64;
  if (grid > cap) grid = cap;
  const uint32_t* vv = reinterpret_cast<const uint32_t*>(d_values_valid);
  const uint32_t* iv = reinterpret_cast<const uint32_t*>(d_indices_valid);
  uint32_t* ov = reinterpret_cast<uint32_t*>(d_out_valid);
#define B2_TAKE_LAUNCH(V, I)                                                                      \
  take_u32_nullable_kernel<V, I><<<(unsigned)grid, kTakeThreads, 0, s>>>(d_values, vv, values_len, \
                                                                         d_indices, iv, idx_len, n, d_out, ov)
  if (vv && iv) B2_TAKE_LAUNCH(true, true);
  else if (vv) B2_TAKE_LAUNCH(true, false);
  else if (iv) B2_TAKE_LAUNCH(false, true);
  else B2_TAKE_LAUNCH(false, false);
#undef B2_TAKE_LAUNCH
  B2_LAUNCH_CHECK(ctx, "take_u32_nullable_kernel");
  return B2_OK;
}

}  // extern "C"
