// microbench_bulk_store.cu — can the copy engine (cp.async.bulk shared -> global, SASS UBLKCP) flush
// the short bucket runs of a radix scatter? Each thread owns runs of S bytes (32 ... 1024) staged in
// shared memory and sends every run to a pseudo-random S-aligned place of a large buffer, either with
// one bulk copy per run (mode "bulk") or the way part_scatter_sectors_kernel stores today: four
// adjacent lanes per 32-byte sector, 8 bytes per lane (mode "st.v2"). Prints ns per run, runs per
// SM clock and GB/s. Run on the B200; not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbb tools/microbench_bulk_store.cu && /tmp/mbb
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kT = 512;
constexpr int kStageBytes = 64 * 1024;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One "tile" = kStageBytes of staged rows = kStageBytes / S runs, dealt round-robin to the threads.
template <int MODE>
__global__ void __launch_bounds__(kT, 2)
k(unsigned char* out, uint64_t out_bytes, int S, int tiles, int group) {
  extern __shared__ __align__(128) unsigned char stage[];
  for (int i = threadIdx.x; i < kStageBytes / 8; i += kT)
    reinterpret_cast<uint2*>(stage)[i] = make_uint2(i, blockIdx.x);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t runs_per_tile = kStageBytes / S;
  const uint64_t slots = out_bytes / S;
  for (int t = 0; t < tiles; ++t) {
    if (MODE == 0) {
      int pending = 0;
      for (uint32_t r = threadIdx.x; r < runs_per_tile; r += kT) {
        const uint64_t slot = mix((blockIdx.x * tiles + t) * runs_per_tile + r) % slots;
        unsigned char* dst = out + slot * S;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"(smem_u32(stage + (uint64_t)r * S)), "r"(S)
                     : "memory");
        if (++pending == group) {
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          pending = 0;
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // the stage may be rewritten once the engine has READ it (what a scatter kernel needs)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
    } else {
      // 8 bytes per lane; consecutive lanes -> consecutive rows of a run
      const uint32_t rows_per_run = S / 8;
      for (uint32_t j = threadIdx.x; j < kStageBytes / 8; j += kT) {
        const uint32_t r = j / rows_per_run, e = j % rows_per_run;
        const uint64_t slot = mix((blockIdx.x * tiles + t) * runs_per_tile + r) % slots;
        const uint2 v = reinterpret_cast<const uint2*>(stage)[j];
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out + slot * S + e * 8), "r"(v.x),
                     "r"(v.y)
                     : "memory");
      }
      __syncthreads();
    }
  }
  if (MODE == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
void run(const char* name, unsigned char* out, uint64_t out_bytes, int S, int group, int nsm, double ghz) {
  const int ctas = nsm * 2, tiles = 64;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
  k<MODE><<<ctas, kT, kStageBytes>>>(out, out_bytes, S, 2, group);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<MODE><<<ctas, kT, kStageBytes>>>(out, out_bytes, S, tiles, group);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const cudaError_t e = cudaGetLastError();
  const double bytes = (double)ctas * tiles * kStageBytes;
  const double runs = bytes / S;
  printf("%-6s S=%4d group=%2d  %8.3f ms  %7.1f GB/s  %6.2f ns/1000 rows  %.3f runs/clk/SM  %s\n", name, S, group,
         best, bytes / best / 1e6, best * 1e6 / (bytes / 8) * 1000, runs / (best * 1e-3 * ghz * 1e9 * nsm),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  const double ghz = p.clockRate / 1e6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, nsm, ghz);
  const uint64_t out_bytes = 8ull << 30;
  unsigned char* out;
  if (cudaMalloc(&out, out_bytes) != cudaSuccess) return 1;
  cudaMemset(out, 0, out_bytes);
  for (int S : {32, 64, 128, 256, 1024}) {
    run<1>("st.v2", out, out_bytes, S, 1, nsm, ghz);
    run<0>("bulk", out, out_bytes, S, 1, nsm, ghz);
    run<0>("bulk", out, out_bytes, S, 8, nsm, ghz);
  }
  cudaFree(out);
  return 0;
}
