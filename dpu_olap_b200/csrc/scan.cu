// scan.cu — single-pass exclusive prefix sum, uint32 counts -> uint64 offsets.
//
// Used by the radix partitioner: the per-(partition, work-unit) histogram, laid out
// partition-major, is turned by ONE flat exclusive scan into the global destination of every
// (partition, unit) run. This replaces the reference's per-DPU prefix_sum
// (dpu/shared/kernels/partition.c:94-137) and the host-side offset bookkeeping under a mutex
// (Partitioner::GetOffsets, host/partition/partitioner.cc:280-312).
// Same chained-scan scheme as the filter: 4096-entry tiles, decoupled look-back.
#include "scan.cuh"

#include "lookback.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

struct ScanWs {
  unsigned long long ticket;
  unsigned long long pad[7];
};

__global__ void __launch_bounds__(kThreads)
exclusive_scan_u32_u64_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out,
                              int64_t n, ScanWs* __restrict__ ws, uint64_t* __restrict__ desc) {
  __shared__ uint64_t warp_tot[kThreads / 32];
  __shared__ int64_t s_tile;
  __shared__ uint64_t s_excl;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (int64_t)atomicAdd(&ws->ticket, 1ull);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t base = tile * kTile + (int64_t)tid * kItems;

  uint32_t v[kItems];
  if (base + kItems <= n) {
    const uint4 a = *reinterpret_cast<const uint4*>(in + base);
    const uint4 b = *reinterpret_cast<const uint4*>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < kItems; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
  }
  uint64_t local = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) local += v[i];
  uint64_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < kThreads / 32 ? warp_tot[lane] : 0;
    uint64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < kThreads / 32) warp_tot[lane] = wi - w;  // exclusive warp offsets
    const uint64_t total = __shfl_sync(0xffffffffu, wi, 31);
    const uint64_t prefix = lookback(desc, tile, total, nullptr);
    if (lane == 0) s_excl = prefix;
  }
  __syncthreads();
  uint64_t run = s_excl + warp_tot[warp] + (incl - local);
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

}  // namespace

size_t b2_scan_ws_bytes(int64_t n) {
  const int64_t ntiles = (n + kTile - 1) / kTile;
  return sizeof(ScanWs) + b2_align_up((size_t)ntiles * 8, 256);
}

int b2_exclusive_scan_u32_u64(b2_ctx* ctx, const uint32_t* d_in, uint64_t* d_out, int64_t n,
                              void* d_ws, size_t ws_bytes, cudaStream_t s) {
  if (n <= 0) return B2_OK;
  const int64_t ntiles = (n + kTile - 1) / kTile;
  if (ws_bytes < b2_scan_ws_bytes(n))
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "scan workspace", "internal sizing error");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 15) == 0, "scan input must be 16 B aligned");
  B2_REQUIRE(ctx, ntiles < (1ll << 31), "scan too large for one launch");
  char* base = static_cast<char*>(d_ws);
  B2_CUDA_OK(ctx, cudaMemsetAsync(base, 0, sizeof(ScanWs) + (size_t)ntiles * 8, s));
  exclusive_scan_u32_u64_kernel<<<(unsigned)ntiles, kThreads, 0, s>>>(
      d_in, d_out, n, reinterpret_cast<ScanWs*>(base),
      reinterpret_cast<uint64_t*>(base + sizeof(ScanWs)));
  B2_LAUNCH_CHECK(ctx, "exclusive_scan_u32_u64_kernel");
  return B2_OK;
}
