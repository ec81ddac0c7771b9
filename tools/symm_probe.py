"""Probe: does torch symmetric memory rendezvous work here, and how fast are peer stores?"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 1 << 27  # 1 GiB of int64
    t = symm_mem.empty(n, dtype=torch.int64, device="cuda")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], file=sys.stderr)
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.int64)
    src = torch.arange(n, dtype=torch.int64, device="cuda")
    for _ in range(2):
        peer.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        peer.copy_(src)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, f"peer copy 1 GiB: {ms:.3f} ms = {n*8/ms/1e6:.1f} GB/s", file=sys.stderr)
    dist.barrier()
    ok = bool((t == src).all())
    print(rank, "data arrived:", ok, file=sys.stderr)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
