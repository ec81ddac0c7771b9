// microbench_scatter_ceiling.cu — the memory-system ceiling of a radix scatter pass, without the
// partitioning arithmetic: every CTA streams its slice of the input through a shared-memory stage and
// writes the stage out as P runs of S bytes to P sequentially advancing streams (what a tile of
// kTile rows looks like after it has been sorted by bucket: run length = tile bytes / P). Compares the
// store paths: "st.v2" (four lanes per 32-byte sector, as part_scatter_sectors_kernel), "bulk" (one
// cp.async.bulk shared -> global per run), "copy" (no scatter at all: the stage goes to one stream).
// Streams are laid out as in the real pass: bucket-major, CTA-minor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_scatter tools/microbench_scatter_ceiling.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: st.v2, 1: bulk per run, 2: plain copy (one stream per CTA)
template <int MODE, int kT>
__global__ void __launch_bounds__(kT, 1)
k(const uint4* __restrict__ in, unsigned char* __restrict__ out, int tile_bytes, int P, int tiles,
  uint64_t stream_bytes) {
  extern __shared__ __align__(128) unsigned char stage[];
  const int S = tile_bytes / P;  // bytes per run
  const uint64_t cta_in = (uint64_t)blockIdx.x * tiles * tile_bytes;
  for (int t = 0; t < tiles; ++t) {
    // ---- stream the tile in (128-bit loads), stage it ----
    const uint4* src = in + (cta_in + (uint64_t)t * tile_bytes) / 16;
    for (int i = threadIdx.x; i < tile_bytes / 16; i += kT) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(src + i));
      reinterpret_cast<uint4*>(stage)[i] = v;
    }
    if (MODE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- write it out as P runs of S bytes: run p -> stream (p, cta) at offset t * S ----
    if (MODE == 0) {
      const int rows_per_run = S / 8;
      for (int j = threadIdx.x; j < tile_bytes / 8; j += kT) {
        const int p = j / rows_per_run, e = j - p * rows_per_run;
        unsigned char* dst = out + ((uint64_t)p * gridDim.x + blockIdx.x) * stream_bytes + (uint64_t)t * S + e * 8;
        const uint2 v = reinterpret_cast<const uint2*>(stage)[j];
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(v.x), "r"(v.y) : "memory");
      }
      __syncthreads();
    } else if (MODE == 1) {
      for (int p = threadIdx.x; p < P; p += kT) {
        unsigned char* dst = out + ((uint64_t)p * gridDim.x + blockIdx.x) * stream_bytes + (uint64_t)t * S;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"(smem_u32(stage + (uint64_t)p * S)), "r"(S)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
    } else {
      unsigned char* dst = out + (uint64_t)blockIdx.x * P * stream_bytes + (uint64_t)t * tile_bytes;
      for (int i = threadIdx.x; i < tile_bytes / 16; i += kT) {
        const uint4 v = reinterpret_cast<const uint4*>(stage)[i];
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + (uint64_t)i * 16), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      }
      __syncthreads();
    }
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, int kT>
void run(const char* name, const uint4* in, unsigned char* out, uint64_t total_bytes, int tile_bytes, int P,
         int ctas_per_sm, int nsm) {
  const int ctas = nsm * ctas_per_sm;
  const int tiles = (int)(total_bytes / ((uint64_t)ctas * tile_bytes));
  const uint64_t stream_bytes = (uint64_t)tiles * (tile_bytes / P);
  cudaFuncSetAttribute(k<MODE, kT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    k<MODE, kT><<<ctas, kT, tile_bytes>>>(in, out, tile_bytes, P, tiles, stream_bytes);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  const cudaError_t e = cudaGetLastError();
  const double bytes = (double)ctas * tiles * tile_bytes;
  printf("%-6s tile=%3d KB x %d CTA/SM (%4d thr)  P=%4d  run=%4d B  %7.3f ms  read+write %7.1f GB/s  %5.2f ns/1000 rows %s\n",
         name, tile_bytes >> 10, ctas_per_sm, kT, P, tile_bytes / P, best, 2 * bytes / best / 1e6,
         best * 1e6 / (bytes / 8) * 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  const uint64_t total = 8ull << 30;  // 2^30 rows of 8 bytes
  uint4* in;
  unsigned char* out;
  if (cudaMalloc(&in, total) != cudaSuccess || cudaMalloc(&out, total + (64 << 20)) != cudaSuccess) return 1;
  cudaMemset(in, 1, total);
  cudaMemset(out, 0, total);
  for (int P : {256, 512, 1024}) {
    run<0, 512>("st.v2", in, out, total, 64 << 10, P, 2, nsm);
    run<1, 512>("bulk", in, out, total, 64 << 10, P, 2, nsm);
    run<0, 1024>("st.v2", in, out, total, 128 << 10, P, 1, nsm);
    run<1, 1024>("bulk", in, out, total, 128 << 10, P, 1, nsm);
    run<1, 512>("bulk", in, out, total, 96 << 10, P, 2, nsm);
  }
  run<2, 512>("copy", in, out, total, 64 << 10, 1024, 2, nsm);
  run<2, 1024>("copy", in, out, total, 128 << 10, 1024, 1, nsm);
  return 0;
}
