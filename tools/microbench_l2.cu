// microbench_l2.cu — whole-GPU rate of random 32-byte accesses into a window that does / does not
// fit the B200's L2 (what an L2-resident hash table can be built from; not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_l2 tools/microbench_l2.cu && tools/mb_l2
// MODE 0: 256-bit ld.global.cg (STRONG.GPU)   1: 256-bit ld.global.nc    2: 128-bit ld.global.cg
// MODE 3: atomicAdd with return (ATOMG)       4: atomicAdd, result unused (REDG)
// MODE 5: 8-byte st.global to a random sector 6: 4-byte st (zeroing one word per sector, sequential)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int MODE, int MLP>
__global__ void __launch_bounds__(256) k(uint32_t* tab, uint32_t nsect, int iters, uint32_t* sink) {
  uint32_t x = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t idx[MLP];
#pragma unroll
    for (int j = 0; j < MLP; ++j) { x = mix(x + j + 1); idx[j] = __umulhi(x, nsect); }
    if (MODE == 0 || MODE == 1) {
      uint32_t r[MLP][8];
#pragma unroll
      for (int j = 0; j < MLP; ++j) {
        const uint32_t* p = tab + 8 * (size_t)idx[j];
        if (MODE == 0)
          asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[j][0]), "=r"(r[j][1]), "=r"(r[j][2]), "=r"(r[j][3]), "=r"(r[j][4]), "=r"(r[j][5]), "=r"(r[j][6]), "=r"(r[j][7]) : "l"(p));
        else
          asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[j][0]), "=r"(r[j][1]), "=r"(r[j][2]), "=r"(r[j][3]), "=r"(r[j][4]), "=r"(r[j][5]), "=r"(r[j][6]), "=r"(r[j][7]) : "l"(p));
      }
#pragma unroll
      for (int j = 0; j < MLP; ++j)
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += r[j][w];
    }
    if (MODE == 2) {
      uint4 r[MLP];
#pragma unroll
      for (int j = 0; j < MLP; ++j) r[j] = __ldcg(reinterpret_cast<const uint4*>(tab + 8 * (size_t)idx[j]));
#pragma unroll
      for (int j = 0; j < MLP; ++j) acc += r[j].x + r[j].y + r[j].z + r[j].w;
    }
    if (MODE == 3) {
      uint32_t r[MLP];
#pragma unroll
      for (int j = 0; j < MLP; ++j) r[j] = atomicAdd(tab + 8 * (size_t)idx[j], 1u);
#pragma unroll
      for (int j = 0; j < MLP; ++j) acc += r[j];
    }
    if (MODE == 4) {
#pragma unroll
      for (int j = 0; j < MLP; ++j) atomicAdd(tab + 8 * (size_t)idx[j], 1u);
    }
    if (MODE == 5) {
#pragma unroll
      for (int j = 0; j < MLP; ++j) *reinterpret_cast<uint2*>(tab + 8 * (size_t)idx[j] + 2) = make_uint2(x, i);
    }
    if (MODE == 6) {
#pragma unroll
      for (int j = 0; j < MLP; ++j) {
        const size_t s = ((size_t)(i * MLP + j) * gridDim.x * 256 + blockIdx.x * 256 + threadIdx.x) % nsect;
        tab[8 * s] = 0;
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int MODE, int MLP>
void run(const char* name, uint32_t* tab, size_t win_bytes, int ctas_per_sm, uint32_t* sink) {
  const uint32_t nsect = (uint32_t)(win_bytes / 32);
  const int grid = 148 * ctas_per_sm, iters = 4096 / MLP;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, MLP><<<grid, 256>>>(tab, nsect, iters / 4, sink);
  cudaEventRecord(e0);
  k<MODE, MLP><<<grid, 256>>>(tab, nsect, iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)grid * 256 * iters * MLP;
  printf("%-28s window %4zu MiB  %d CTAs/SM x256 thr, MLP %d: %7.1f G ops/s  (%6.2f ms)  err=%d\n", name,
         win_bytes >> 20, ctas_per_sm, MLP, ops / ms * 1e-6, ms, (int)cudaGetLastError());
}

int main() {
  uint32_t *tab, *sink;
  const size_t maxb = (size_t)1 << 30;
  cudaMalloc(&tab, maxb);
  cudaMalloc(&sink, 4);
  cudaMemset(tab, 0, maxb);
  for (size_t mib : {8, 32, 48, 64, 96, 256, 1024}) {
    const size_t w = mib << 20;
    run<0, 4>("ld.cg 256-bit", tab, w, 4, sink);
    run<0, 8>("ld.cg 256-bit", tab, w, 2, sink);
    run<1, 4>("ld.nc 256-bit", tab, w, 4, sink);
    run<2, 8>("ld.cg 128-bit", tab, w, 4, sink);
    run<3, 8>("atomicAdd returning", tab, w, 4, sink);
    run<4, 8>("atomicAdd (red)", tab, w, 4, sink);
    run<5, 8>("st 8 B random sector", tab, w, 4, sink);
    run<6, 8>("st 4 B per sector, sequential", tab, w, 4, sink);
  }
  return 0;
}
