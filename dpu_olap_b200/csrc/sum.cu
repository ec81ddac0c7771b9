// sum.cu — Sum aggregate of a uint32 column into uint64.
//
// Replaces the reference's DPU aggregate program: per-tasklet partial sums over 256-element
// blocks (dpu/shared/kernels/aggr.c:16-33, `sum` dpu/aggr/main.c:44-51), tasklet 0 adding the 16
// partials (main.c:75-89) and the host adding one partial per DPU (host/aggr/aggr_dpu.cc:82-84).
//
// B200 design: the op is a pure HBM read stream (4 B/row). One persistent grid of
// 2 CTAs x 512 threads per SM; every thread keeps 8 independent 128-bit streaming loads in flight
// (128 KB per SM outstanding), accumulates in 64-bit registers, reduces by warp shuffle, and the
// last CTA to finish (atomic ticket) folds the per-CTA partials — one launch, no memset, and the
// result is written (not accumulated), so the call is idempotent.
#include "common.cuh"

namespace {

constexpr int kSumThreads = 512;
constexpr int kSumUnroll = 8;               // uint4 loads in flight per thread
constexpr int kSumMaxCtas = 448;            // CTAs per launch (2 per SM; B200 has 148 SMs)
// Every launch gets its OWN partials and ticket: ctx->d_small is a ring of kSumSlots scratch slots
// (sums | counts | ticket), claimed round-robin on the host, so launches of one ctx that are in
// flight on different streams never share state (up to kSumSlots of them at a time).
constexpr int kSumSlots = 16;
constexpr size_t kSumSlotBytes = 2 * kSumMaxCtas * 8 + 1024;  // 8 KB; the ticket sits after sums and counts
constexpr size_t kTicketOffset = 2 * kSumMaxCtas * 8;
static_assert(kSumSlots * kSumSlotBytes <= 128 * 1024, "b2_ctx_create allocates 128 KB of small scratch");

__device__ __forceinline__ uint64_t sum4(uint4 v) {
  return ((uint64_t)v.x + v.y) + ((uint64_t)v.z + v.w);
}
// Fused filter -> sum: rows that fail `v < thr` contribute 0 to the sum and to the count.
// acc: sum in the value lanes; cnt: number of selected rows.
__device__ __forceinline__ void sum4_lt(uint4 v, uint32_t thr, uint64_t& acc, uint32_t& cnt) {
  const uint32_t a = v.x < thr, b = v.y < thr, c = v.z < thr, d = v.w < thr;
  acc += ((uint64_t)(a ? v.x : 0u) + (b ? v.y : 0u)) + ((uint64_t)(c ? v.z : 0u) + (d ? v.w : 0u));
  cnt += a + b + c + d;
}

// kFiltered: the fused pipeline filter(v < thr) -> sum the reference leaves commented out in its
// Acero plan (host/aggr/aggr_native.cc:59-65): one read of the column, nothing materialised; the
// per-CTA partials then carry (sum, count) pairs.
template <bool kFiltered>
__global__ void __launch_bounds__(kSumThreads, 2)
sum_u32_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t thr, uint64_t* __restrict__ partials,
               unsigned int* __restrict__ ticket, uint64_t* __restrict__ out,
               uint64_t* __restrict__ out_count) {
  // Split [0,n) into a scalar head (to reach 16 B alignment), a vector body and a scalar tail.
  const uintptr_t addr = reinterpret_cast<uintptr_t>(in);
  int64_t head = (int64_t)(((16 - (addr & 15)) & 15) >> 2);
  if (head > n) head = n;
  const int64_t nvec = (n - head) >> 2;
  const int64_t tail_start = head + (nvec << 2);
  const uint4* __restrict__ vin = reinterpret_cast<const uint4*>(in + head);

  uint64_t acc = 0;
  uint32_t cnt32 = 0;   // selected rows of this thread since the last flush
  uint64_t cnt = 0;
  const int64_t chunk = (int64_t)kSumThreads * kSumUnroll;  // uint4 per CTA iteration
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
    uint4 v[kSumUnroll];
    if (base + chunk <= nvec) {
#pragma unroll
      for (int u = 0; u < kSumUnroll; ++u) v[u] = ld_stream_v4(vin + base + u * kSumThreads + threadIdx.x);
#pragma unroll
      for (int u = 0; u < kSumUnroll; ++u) {
        if (kFiltered) sum4_lt(v[u], thr, acc, cnt32);
        else acc += sum4(v[u]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kSumUnroll; ++u) {
        const int64_t i = base + u * kSumThreads + threadIdx.x;
        if (i < nvec) {
          const uint4 q = ld_stream_v4(vin + i);
          if (kFiltered) sum4_lt(q, thr, acc, cnt32);
          else acc += sum4(q);
        }
      }
    }
    if (kFiltered) {  // 32 rows per iteration at most: flush long before the 32-bit counter could wrap
      cnt += cnt32;
      cnt32 = 0;
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = threadIdx.x; i < head; i += kSumThreads)
      if (!kFiltered || in[i] < thr) { acc += in[i]; ++cnt; }
    for (int64_t i = tail_start + threadIdx.x; i < n; i += kSumThreads)
      if (!kFiltered || in[i] < thr) { acc += in[i]; ++cnt; }
  }

  __shared__ uint64_t warp_sums[kSumThreads / 32];
  __shared__ uint64_t warp_cnts[kSumThreads / 32];
  __shared__ bool is_last;
  acc = warp_reduce_sum_u64(acc);
  if (kFiltered) cnt = warp_reduce_sum_u64(cnt);
  if (lane_id() == 0) {
    warp_sums[threadIdx.x >> 5] = acc;
    if (kFiltered) warp_cnts[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint64_t v = threadIdx.x < kSumThreads / 32 ? warp_sums[threadIdx.x] : 0;
    v = warp_reduce_sum_u64(v);
    uint64_t c = 0;
    if (kFiltered) {
      c = threadIdx.x < kSumThreads / 32 ? warp_cnts[threadIdx.x] : 0;
      c = warp_reduce_sum_u64(c);
    }
    if (threadIdx.x == 0) {
      partials[blockIdx.x] = v;
      if (kFiltered) partials[kSumMaxCtas + blockIdx.x] = c;
      __threadfence();
      const unsigned int t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    uint64_t v = 0, c = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kSumThreads) {
      v += *reinterpret_cast<volatile uint64_t*>(partials + i);
      if (kFiltered) c += *reinterpret_cast<volatile uint64_t*>(partials + kSumMaxCtas + i);
    }
    v = warp_reduce_sum_u64(v);
    if (kFiltered) c = warp_reduce_sum_u64(c);
    __syncthreads();
    if (lane_id() == 0) {
      warp_sums[threadIdx.x >> 5] = v;
      if (kFiltered) warp_cnts[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      uint64_t w = threadIdx.x < kSumThreads / 32 ? warp_sums[threadIdx.x] : 0;
      w = warp_reduce_sum_u64(w);
      uint64_t wc = 0;
      if (kFiltered) {
        wc = threadIdx.x < kSumThreads / 32 ? warp_cnts[threadIdx.x] : 0;
        wc = warp_reduce_sum_u64(wc);
      }
      if (threadIdx.x == 0) {
        *out = w;
        if (kFiltered && out_count) *out_count = wc;
        *ticket = 0;  // re-arm for the next launch on this ctx
      }
    }
  }
}

}  // namespace

namespace {
int sum_launch(b2_ctx* ctx, const uint32_t* d_in, int64_t n, bool filtered, uint32_t thr, uint64_t* d_sum,
               uint64_t* d_count, cudaStream_t s) {
  B2_REQUIRE(ctx, n >= 0, "n must be >= 0");
  B2_REQUIRE(ctx, d_sum != nullptr, "d_sum is null");
  B2_REQUIRE(ctx, n == 0 || d_in != nullptr, "d_in is null");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 3) == 0, "d_in must be 4-byte aligned");
  const int64_t chunk_elems = (int64_t)kSumThreads * kSumUnroll * 4;
  int64_t want = (n + chunk_elems - 1) / chunk_elems;
  int grid = ctx->sm_count * 2;
  if (want < grid) grid = want > 0 ? (int)want : 1;
  if (grid > kSumMaxCtas) grid = kSumMaxCtas;
  char* slot = static_cast<char*>(ctx->d_small) + (size_t)(ctx->sum_slot++ % kSumSlots) * kSumSlotBytes;
  uint64_t* partials = reinterpret_cast<uint64_t*>(slot);  // [kSumMaxCtas] sums | [kSumMaxCtas] counts
  unsigned int* ticket = reinterpret_cast<unsigned int*>(slot + kTicketOffset);
  if (filtered)
    sum_u32_kernel<true><<<grid, kSumThreads, 0, s>>>(d_in, n, thr, partials, ticket, d_sum, d_count);
  else
    sum_u32_kernel<false><<<grid, kSumThreads, 0, s>>>(d_in, n, 0u, partials, ticket, d_sum, nullptr);
  B2_LAUNCH_CHECK(ctx, "sum_u32_kernel");
  return B2_OK;
}
}  // namespace

extern "C" int b2_sum_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t n, uint64_t* d_sum,
                              void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  return sum_launch(ctx, d_in, n, false, 0u, d_sum, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int b2_sum_lt_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t n, uint32_t threshold,
                                 uint64_t* d_sum, uint64_t* d_count, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  return sum_launch(ctx, d_in, n, true, threshold, d_sum, d_count, static_cast<cudaStream_t>(stream));
}
