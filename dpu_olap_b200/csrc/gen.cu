// gen.cu — device-side synthetic inputs, bit-identical to the reference's generator.
//
// The reference fills every uint32 array with (host/generator/random.cc:103-109)
//     pcg32_fast rng(seed_++);  std::uniform_int_distribution<uint32_t> dist(min_, max_);
//     std::generate(data, data + n, [&] { return dist(rng); });
// pcg32_fast is PCG's mcg_xsh_rs_64_32: a 64-bit multiplicative congruential generator
//     state0 = seed | 3;  out_i = xsh_rs(state_i);  state_{i+1} = state_i * 6364136223846793005
// with xsh_rs(s) = ((s >> 22) ^ s) >> (22 + (s >> 61)) truncated to 32 bits. Because there is no
// additive term, state_i = state0 * M^i (mod 2^64): every thread jumps straight to its first
// element with a square-and-multiply power and then strides by a constant M^stride, so a
// 64 GiB column (SF=2048 filter input) is generated at HBM write speed instead of being staged
// through host RAM. libstdc++'s uniform_int_distribution<uint32_t> over a 32-bit URNG is Lemire's
// multiply-shift: for a span of 2^k it never rejects and returns lo + (g >> (32 - k)); for the
// full range it returns g itself. Those are the only spans the reference's fixtures use
// (2^32 for data, 2^21 for fk, 2^22 / 2^16 for take indices); other spans reject draws, which
// breaks random access, and are refused with B2_ERR_UNSUPPORTED.
#include <vector>

#include "common.cuh"

namespace {

constexpr uint64_t kPcgMult = 6364136223846793005ull;
constexpr int kThreads = 256;
constexpr int64_t kChunk = 65536;  // elements per CTA

struct GenParam {   // one per batch
  uint64_t state0;  // seed | 3
  uint32_t lo;
  uint32_t bits;    // k: span = 2^k, 32 = full range
};

__host__ __device__ inline uint64_t mulpow(uint64_t base, uint64_t e) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r *= base;
    base *= base;
    e >>= 1;
  }
  return r;
}

__device__ __forceinline__ uint32_t pcg_xsh_rs(uint64_t s) {
  return (uint32_t)(((s >> 22) ^ s) >> (22 + (uint32_t)(s >> 61)));
}

__global__ void __launch_bounds__(kThreads)
gen_u32_kernel(const GenParam* __restrict__ params, int64_t batch_len, int64_t chunks_per_batch,
               uint64_t stride_mult /* M^kThreads */, uint32_t* __restrict__ out) {
  const int64_t b = blockIdx.x / chunks_per_batch;
  const int64_t c = blockIdx.x - b * chunks_per_batch;
  const GenParam p = params[b];
  const int64_t begin = c * kChunk;
  int64_t end = begin + kChunk;
  if (end > batch_len) end = batch_len;
  uint64_t s = p.state0 * mulpow(kPcgMult, (uint64_t)(begin + threadIdx.x));
  uint32_t* __restrict__ dst = out + b * batch_len;
  for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) {
    const uint32_t g = pcg_xsh_rs(s);
    s *= stride_mult;
    dst[i] = p.lo + (uint32_t)(((uint64_t)g << p.bits) >> 32);
  }
}

__global__ void __launch_bounds__(kThreads)
iota_u32_kernel(uint64_t start, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kThreads)
    out[i] = (uint32_t)(start + (uint64_t)i);
}

}  // namespace

int b2_ctx_reserve_ws(b2_ctx* ctx, size_t bytes) {
  if (ctx->ws_bytes >= bytes) return B2_OK;
  if (ctx->d_ws) {
    cudaDeviceSynchronize();
    cudaFree(ctx->d_ws);
    ctx->d_ws = nullptr;
    ctx->ws_bytes = 0;
  }
  bytes = b2_align_up(bytes, 1 << 20);
  B2_CUDA_OK(ctx, cudaMalloc(&ctx->d_ws, bytes));
  ctx->ws_bytes = bytes;
  return B2_OK;
}

extern "C" {

int b2_gen_u32_dev(b2_ctx* ctx, const uint64_t* data_seeds, const uint32_t* lo, const uint32_t* hi,
                   int64_t nbatches, int64_t batch_len, uint32_t* d_out, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, nbatches >= 0 && batch_len >= 0, "negative size");
  if (nbatches * batch_len == 0) return B2_OK;
  B2_REQUIRE(ctx, data_seeds && d_out, "null pointer");
  B2_REQUIRE(ctx, (lo == nullptr) == (hi == nullptr), "lo and hi must both be given or both NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<GenParam> params((size_t)nbatches);
  for (int64_t b = 0; b < nbatches; ++b) {
    GenParam p;
    p.state0 = data_seeds[b] | 3ull;
    p.lo = 0;
    p.bits = 32;
    if (lo) {
      if (hi[b] < lo[b]) return b2_set_error(ctx, B2_ERR_INVALID, "gen", "hi < lo");
      const uint64_t span = (uint64_t)hi[b] - lo[b] + 1;
      if (span & (span - 1))
        return b2_set_error(ctx, B2_ERR_UNSUPPORTED, "gen",
                            "span hi-lo+1 must be a power of two (rejection sampling otherwise)");
      p.lo = lo[b];
      p.bits = 0;
      while ((1ull << p.bits) < span) ++p.bits;
    }
    params[(size_t)b] = p;
  }
  const size_t bytes = (size_t)nbatches * sizeof(GenParam);
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  B2_RETURN_NOT_OK(b2_ctx_reserve_ws(ctx, bytes));
  B2_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_ws, params.data(), bytes, cudaMemcpyHostToDevice, s));
  const int64_t cpb = (batch_len + kChunk - 1) / kChunk;
  const int64_t grid = nbatches * cpb;
  B2_REQUIRE(ctx, grid < (1ll << 31), "too many chunks for one launch");
  gen_u32_kernel<<<(unsigned)grid, kThreads, 0, s>>>(static_cast<const GenParam*>(ctx->d_ws),
                                                     batch_len, cpb, mulpow(kPcgMult, kThreads),
                                                     d_out);
  B2_LAUNCH_CHECK(ctx, "gen_u32_kernel");
  // params live in ctx->d_ws, which the next call may overwrite
  B2_CUDA_OK(ctx, cudaStreamSynchronize(s));
  return B2_OK;
}

int b2_iota_u32_dev(b2_ctx* ctx, uint64_t start, int64_t n, uint32_t* d_out, void* stream) {
  if (!ctx) return B2_ERR_INVALID;
  b2_device_scope dev_scope(ctx);
  B2_REQUIRE(ctx, n >= 0, "negative size");
  if (n == 0) return B2_OK;
  B2_REQUIRE(ctx, d_out != nullptr, "null pointer");
  int64_t grid = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (grid > cap) grid = cap;
  iota_u32_kernel<<<(unsigned)grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(start, n, d_out);
  B2_LAUNCH_CHECK(ctx, "iota_u32_kernel");
  return B2_OK;
}

}  // extern "C"
