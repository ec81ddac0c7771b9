"""Microbenchmark (2 GPUs): NVLink peer-store bandwidth as a function of run length and alignment.
Each thread block writes `run` consecutive 8-byte elements at pseudo-random run-granular positions
of the peer buffer, the pattern the fused shuffle produces (run = rows of one bucket per tile)."""
import os, sys
import torch, torch.distributed as dist
from torch.utils.cpp_extension import load_inline

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import torch.distributed._symmetric_memory as symm_mem

src = r'''
#include <torch/extension.h>
#include <cuda_runtime.h>
__global__ void scatter_runs(const unsigned long long* __restrict__ in, unsigned long long* out, long n, int run, int shift, long nruns) {
  // thread i copies element i of the input to a permuted run position: run r -> (r * 2654435761) % nruns
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long r = i / run, k = i - r * run;
    long pr = (long)(((unsigned long long)r * 2654435761ull) % (unsigned long long)nruns);
    long dst = pr * run + k + shift;
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" :: "l"(out + dst), "l"(in[i]) : "memory");
  }
}
void launch(torch::Tensor in, long out_ptr, long n, int run, int shift) {
  long nruns = n / run;
  scatter_runs<<<148 * 8, 512>>>((const unsigned long long*)in.data_ptr(), (unsigned long long*)out_ptr, nruns * run, run, shift, nruns);
}
'''
mod = load_inline(name=f"p2pprobe{rank}", cpp_sources="void launch(torch::Tensor in, long out_ptr, long n, int run, int shift);",
                  cuda_sources=src, functions=["launch"], extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=False)
n = 1 << 27  # 1 GiB
t = symm_mem.empty(n + 64, dtype=torch.int64, device="cuda")
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
peer_ptr = hdl.buffer_ptrs[(rank + 1) % world]
local_ptr = hdl.buffer_ptrs[rank]
inp = torch.arange(n, dtype=torch.int64, device="cuda")
for target, name in ((peer_ptr, "peer"), (local_ptr, "local")):
    for run in (8, 16, 32, 64, 512, 4096):
        for shift in (0, 1):
            for _ in range(2):
                mod.launch(inp, target, n, run, shift)
            torch.cuda.synchronize(); dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                mod.launch(inp, target, n, run, shift)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            if rank == 0:
                print(f"{name} run={run*8:6d} B shift={shift*8} B: {n*8/ms/1e6:7.1f} GB/s", file=sys.stderr, flush=True)
            dist.barrier()
dist.destroy_process_group()
