// filter.cu — order-preserving selection `v < threshold` over a uint32 column (stream compaction).
//
// Replaces the reference's DPU filter program (dpu/shared/kernels/filter.c:57-177): there each of
// 16 tasklets compacts a 128-element block in WRAM and a serial handshake chain
// (handshake_sync, filter.c:28-55) hands the running output offset from tasklet to tasklet, with
// a `write_carry` fix-up for 8-byte MRAM alignment (:109-131) and two barriers per round. The
// predicate is `item < (1 << 30)` (filter.c:25; Acero side: filter_native.cc:59).
//
// B200 design: ONE pass over the column (4 B read per row + 4 B written per selected row):
//   * a tile is 4096 rows = 256 threads x 4 x 128-bit streaming loads, laid out so every warp
//     owns 512 consecutive rows (each warp-level load is one fully coalesced 512 B request);
//     six CTAs are resident per SM so the phases of different tiles overlap;
//   * ranks inside a warp come from __ballot_sync/__popc (no shuffles); the 32 (warp,segment)
//     counts of the CTA are scanned redundantly by every warp (one count per lane);
//   * tiles are chained by a decoupled look-back over single-word 64-bit descriptors
//     (2 status bits | 62-bit running count, so 2^34-row columns need no second level), run by
//     warp 0 WHILE the other warps already compact the tile into shared memory — tile-local
//     positions do not depend on the running count;
//   * selected rows are staged in shared memory and written with fully coalesced stores.
// Tiles never straddle a batch boundary, so the inclusive count of the last tile of batch b is
// the end offset of result chunk b (FilterDpu::GetResult returns one chunk per input batch,
// host/filter/filter_dpu.cc:89-96,162-166) — a tiny second kernel collects those.
#include "common.cuh"
#include "lookback.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                       // uint4 loads per thread
constexpr int kTile = kThreads * kVecPerThread * 4;    // 4096 rows
constexpr int kWarps = kThreads / 32;
constexpr int kSegs = kWarps * kVecPerThread;          // 32 (warp, segment) counts per tile
static_assert(kSegs == 32, "one (warp, segment) count per lane");

struct FilterWs {           // header of the caller workspace (reserved)
  unsigned long long pad[8];
};

static inline int64_t tiles_of(int64_t len) { return (len + kTile - 1) / kTile; }

// 1 if a < b (unsigned) else 0, as an INTEGER: keeps the 16 per-row predicate results of a thread
// in one mask register instead of 16 predicate registers (ptxas has 7 and gives up otherwise).
__device__ __forceinline__ uint32_t lt_bit(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("set.lt.u32.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r & 1u;
}

// Tile id = blockIdx.x: CTAs are dispatched in increasing block order, so every predecessor of a
// tile that waits in the look-back is already resident or finished (the same assumption CUB's
// single-pass scan and select make).
template <bool kUniform>
__global__ void __launch_bounds__(kThreads, 6)
filter_lt_u32_kernel(const uint32_t* __restrict__ in, uint32_t thr, uint32_t* __restrict__ out,
                     int64_t batch_len, int64_t tiles_per_batch,          // uniform layout
                     const int64_t* __restrict__ batch_off,               // ragged layout
                     const int64_t* __restrict__ tile_first, int64_t nbatches,
                     const int64_t* __restrict__ carry_in, uint64_t* __restrict__ desc) {
  __shared__ uint32_t stage[kTile];
  __shared__ uint32_t seg_cnt[kSegs];
  __shared__ uint64_t s_excl;
  __shared__ uint32_t seg_off[kSegs];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;

  // tile -> rows [row0, row0 + len)
  int64_t row0, len;
  if (kUniform) {
    const int64_t b = tile / tiles_per_batch, k = tile - b * tiles_per_batch;
    row0 = b * batch_len + k * kTile;
    len = batch_len - k * kTile;
  } else {
    int64_t lo = 0, hi = nbatches;  // last b with tile_first[b] <= tile
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (tile_first[mid] <= tile) lo = mid; else hi = mid;
    }
    const int64_t k = tile - tile_first[lo];
    row0 = batch_off[lo] + k * kTile;
    len = batch_off[lo + 1] - row0;
  }
  if (len > kTile) len = kTile;

  // ---- load + predicate: warp w owns rows [w*512, w*512+512), segment j = 128 rows ----
  uint32_t v[kVecPerThread][4];
  uint32_t mask = 0;  // bit (j*4+e) set <=> element selected
  const uint32_t* __restrict__ src = in + row0;
  const uint32_t e0 = warp * (kTile / kWarps) + lane * 4;  // + j*128 + e
  if (len == kTile && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* __restrict__ vsrc = reinterpret_cast<const uint4*>(src);
    uint4 q[kVecPerThread];
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) q[j] = ld_stream_v4(vsrc + ((e0 + j * 128) >> 2));
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      v[j][0] = q[j].x; v[j][1] = q[j].y; v[j][2] = q[j].z; v[j][3] = q[j].w;
#pragma unroll
      for (int e = 0; e < 4; ++e) mask |= lt_bit(v[j][e], thr) << (j * 4 + e);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t i = e0 + j * 128 + e;
        v[j][e] = 0;
        if ((int64_t)i < len) {
          v[j][e] = ld_stream_u32(src + i);
          mask |= lt_bit(v[j][e], thr) << (j * 4 + e);
        }
      }
    }
  }

  // ---- ranks inside the warp: order is (segment j, lane, element e) ----
  uint32_t lane_excl[kVecPerThread];
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < kVecPerThread; ++j) {
    uint32_t below = 0, total = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t b = __ballot_sync(0xffffffffu, (mask >> (j * 4 + e)) & 1u);
      below += __popc(b & lt);
      total += __popc(b);
    }
    lane_excl[j] = below;
    if (lane == 0) seg_cnt[warp * kVecPerThread + j] = total;
  }
  __syncthreads();

  // ---- every warp scans the 32 segment counts itself (one per lane; warp 0 needs the total) ----
  const uint32_t c = seg_cnt[lane];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  // (offsets go through shared memory: fetching them with a shuffle from a warp-uniform lane
  //  index makes ptxas 12.9 fail register allocation for this kernel)
  if (warp == 1) seg_off[lane] = incl - c;
  __syncthreads();
  // warp 0 chains the tile with its predecessors while the other warps already compact
  if (warp == 0) {
    const uint64_t prefix = lookback(desc, tile, total, carry_in);
    if (lane == 0) s_excl = prefix;
  }

  // ---- compact into shared memory (positions are tile-local: no dependence on the look-back) ----
#pragma unroll
  for (int j = 0; j < kVecPerThread; ++j) {
    uint32_t p = seg_off[warp * kVecPerThread + j] + lane_excl[j];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if ((mask >> (j * 4 + e)) & 1u) stage[p++] = v[j][e];
    }
  }
  __syncthreads();

  // ---- coalesced write-out ----
  uint32_t* __restrict__ dst = out + s_excl;
  for (uint32_t i = tid; i < total; i += kThreads) st_stream_u32(dst + i, stage[i]);
}

// batch_end[b] = inclusive count at the last tile of batch b (carried over empty batches).
__global__ void filter_batch_end_kernel(const uint64_t* __restrict__ desc, int64_t nbatches,
                                        int64_t tiles_per_batch,
                                        const int64_t* __restrict__ tile_first,
                                        const int64_t* __restrict__ carry_in,
                                        int64_t* __restrict__ batch_end,
                                        int64_t* __restrict__ total) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t carry = carry_in ? *carry_in : 0;
  if (b < nbatches) {
    const int64_t last = tile_first ? tile_first[b + 1] : (b + 1) * tiles_per_batch;  // exclusive
    batch_end[b] = last > 0 ? (int64_t)(desc[last - 1] & kValMask) : carry;
  }
  if (b == 0 && total) {
    const int64_t ntiles = tile_first ? tile_first[nbatches] : nbatches * tiles_per_batch;
    *total = ntiles > 0 ? (int64_t)(desc[ntiles - 1] & kValMask) : carry;
  }
}

int filter_launch(b2_ctx* ctx, const uint32_t* d_in, int64_t nbatches, int64_t batch_len,
                  const int64_t* h_batch_off, const int64_t* d_batch_off, uint32_t thr,
                  uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                  const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, cudaStream_t s) {
  const bool uniform = (h_batch_off == nullptr);
  int64_t ntiles = 0;
  if (uniform) {
    ntiles = nbatches * tiles_of(batch_len);
  } else {
    for (int64_t b = 0; b < nbatches; ++b) ntiles += tiles_of(h_batch_off[b + 1] - h_batch_off[b]);
  }
  const size_t desc_bytes = b2_align_up((size_t)ntiles * 8, 256);
  const size_t tf_bytes = uniform ? 0 : b2_align_up((size_t)(nbatches + 1) * 8, 256);
  const size_t need = sizeof(FilterWs) + desc_bytes + tf_bytes;
  if (ws_bytes < need || (need > 0 && d_ws == nullptr))
    return b2_set_error(ctx, B2_ERR_WORKSPACE, "filter workspace", "use b2_filter_ws_bytes()");
  B2_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0, "workspace must be 16 B aligned");
  char* base = static_cast<char*>(d_ws);
  uint64_t* desc = reinterpret_cast<uint64_t*>(base + sizeof(FilterWs));
  int64_t* tile_first = uniform ? nullptr : reinterpret_cast<int64_t*>(base + sizeof(FilterWs) + desc_bytes);

  // zero every descriptor (status 0 = not published)
  B2_CUDA_OK(ctx, cudaMemsetAsync(base, 0, sizeof(FilterWs) + desc_bytes, s));
  if (!uniform) {
    // tile_first is tiny; build it on the host and ship it (pageable copy is staged before return)
    std::string buf((size_t)(nbatches + 1) * 8, '\0');
    int64_t* tf = reinterpret_cast<int64_t*>(&buf[0]);
    tf[0] = 0;
    for (int64_t b = 0; b < nbatches; ++b) tf[b + 1] = tf[b] + tiles_of(h_batch_off[b + 1] - h_batch_off[b]);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(tile_first, tf, (size_t)(nbatches + 1) * 8, cudaMemcpyHostToDevice, s));
  }
  if (ntiles > 0) {
    B2_REQUIRE(ctx, ntiles < (1ll << 31), "too many tiles for one launch");
    if (uniform) {
      filter_lt_u32_kernel<true><<<(unsigned)ntiles, kThreads, 0, s>>>(
          d_in, thr, d_out, batch_len, tiles_of(batch_len), nullptr, nullptr, nbatches, d_carry_in,
          desc);
    } else {
      filter_lt_u32_kernel<false><<<(unsigned)ntiles, kThreads, 0, s>>>(
          d_in, thr, d_out, 0, 0, d_batch_off, tile_first, nbatches, d_carry_in, desc);
    }
    B2_LAUNCH_CHECK(ctx, "filter_lt_u32_kernel");
  }
  if (nbatches > 0 || d_total) {
    const int64_t nb = nbatches > 0 ? nbatches : 1;
    filter_batch_end_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(
        desc, nbatches, tiles_of(batch_len), tile_first, d_carry_in, d_batch_end, d_total);
    B2_LAUNCH_CHECK(ctx, "filter_batch_end_kernel");
  }
  return B2_OK;
}

}  // namespace

extern "C" {

size_t b2_filter_ws_bytes(int64_t nbatches, int64_t batch_len) {
  if (nbatches < 0 || batch_len < 0) return 0;
  return sizeof(FilterWs) + b2_align_up((size_t)(nbatches * tiles_of(batch_len)) * 8, 256);
}

size_t b2_filter_ragged_ws_bytes(const int64_t* h_batch_off, int64_t nbatches) {
  if (nbatches < 0 || (nbatches > 0 && !h_batch_off)) return 0;
  int64_t ntiles = 0;
  for (int64_t b = 0; b < nbatches; ++b) ntiles += tiles_of(h_batch_off[b + 1] - h_batch_off[b]);
  return sizeof(FilterWs) + b2_align_up((size_t)ntiles * 8, 256) +
         b2_align_up((size_t)(nbatches + 1) * 8, 256);
}

int b2_filter_lt_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t nbatches, int64_t batch_len,
                         uint32_t threshold, uint32_t* d_out, int64_t* d_batch_end,
                         int64_t* d_total, const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                         void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  B2_REQUIRE(ctx, nbatches >= 0 && batch_len >= 0, "negative size");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  B2_REQUIRE(ctx, nbatches * batch_len == 0 || (d_in && d_out), "null column pointer");
  return filter_launch(ctx, d_in, nbatches, batch_len, nullptr, nullptr, threshold, d_out,
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

int b2_filter_lt_u32_ragged_dev(b2_ctx* ctx, const uint32_t* d_in, const int64_t* h_batch_off,
                                const int64_t* d_batch_off, int64_t nbatches, uint32_t threshold,
                                uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                                const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                                void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  B2_REQUIRE(ctx, nbatches >= 0, "negative size");
  B2_REQUIRE(ctx, h_batch_off && d_batch_off, "batch offset tables are null");
  for (int64_t b = 0; b < nbatches; ++b)
    B2_REQUIRE(ctx, h_batch_off[b + 1] >= h_batch_off[b], "batch offsets must be non-decreasing");
  B2_REQUIRE(ctx, nbatches == 0 || d_batch_end != nullptr, "d_batch_end is null");
  return filter_launch(ctx, d_in, nbatches, 0, h_batch_off, d_batch_off, threshold, d_out,
                       d_batch_end, d_total, d_carry_in, d_ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

}  // extern "C"
