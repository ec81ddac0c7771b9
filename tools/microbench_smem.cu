// microbench_smem.cu — per-SM throughput of the shared-memory primitives the partitioner/join can
// be built from (run on the B200; not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_smem.cu && /tmp/mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t wang(uint32_t key) {
  key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
  key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16); return key;
}

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, long long* cycles) {
  __shared__ uint32_t h[8192];
  for (int i = threadIdx.x; i < 8192; i += 512) h[i] = 0;
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = wang(x + i);
    const uint32_t b = x >> 22;  // 1024 bins
    if (MODE == 0) { atomicAdd(&h[b], 1u); }                       // RED-style (result unused)
    if (MODE == 1) { acc += atomicAdd(&h[b], 1u); }                // ATOMS with return
    if (MODE == 2) { acc += __match_any_sync(0xffffffffu, b); }    // MATCH.ANY
    if (MODE == 3) { acc += atomicCAS(&h[x >> 19], 0u, x | 1u); }  // CAS on 8192 slots
    if (MODE == 4) { acc += h[b]; }                                // plain LDS, random banks
    if (MODE == 5) { h[b] = x; }                                   // plain STS, random banks
    if (MODE == 6) { acc += b; }                                   // hash only (ALU floor)
    if (MODE == 7) {                                               // match + leader RMW + shfl (rank)
      const uint32_t peers = __match_any_sync(0xffffffffu, b);
      const int leader = __ffs(peers) - 1;
      uint32_t before = 0;
      if ((threadIdx.x & 31) == leader) { before = h[b]; h[b] = before + __popc(peers); }
      acc += __shfl_sync(0xffffffffu, before, leader);
      __syncwarp();
    }
    if (MODE == 8) {                                               // warp-private u16 atomics emulation: 32-bit atomicAdd on private region
      acc += atomicAdd(&h[((threadIdx.x >> 5) & 7) * 1024 + b], 1u);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 512 + threadIdx.x] = acc + h[threadIdx.x];
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  int iters = 4096, nsm = 148;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4 * 512 * nsm * 4);
  cudaMalloc(&cyc, 8 * nsm * 4);
  k<MODE><<<nsm * ctas_per_sm, 512>>>(out, 16, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<nsm * ctas_per_sm, 512>>>(out, iters, cyc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[4 * 148]; cudaMemcpy(h, cyc, 8 * nsm * ctas_per_sm, cudaMemcpyDeviceToHost);
  double warp_instr_per_sm = (double)iters * 16 * ctas_per_sm;
  printf("%-44s ctas/SM=%d  %8.3f ms  cycles(cta0)=%lld  => %.2f cyc per warp-op per SM, %.1f Glane-ops/s chip\n",
         name, ctas_per_sm, ms, h[0], (double)h[0] / warp_instr_per_sm,
         (double)iters * 512 * nsm * ctas_per_sm / (ms * 1e-3) / 1e9);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<6>("hash only (ALU floor)", c);
    run<0>("smem atomicAdd, result unused, 1024 bins", c);
    run<1>("smem atomicAdd with return, 1024 bins", c);
    run<8>("smem atomicAdd return, warp-private bins", c);
    run<2>("__match_any_sync", c);
    run<7>("match_any + leader LDS/STS + shfl (rank)", c);
    run<3>("smem atomicCAS, 8192 slots", c);
    run<4>("LDS random", c);
    run<5>("STS random", c);
  }
  return 0;
}
