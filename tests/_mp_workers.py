"""Worker processes of the real multi-GPU tests (spawned by tests/test_gpu_multi.py, one per GPU)."""
from __future__ import annotations

import os
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def p2p_join_worker(rank: int, world: int, port: int, nb_total: int, batch: int, out_dir: str, dup: bool,
                    shares: int | None = None):
    """One rank of P2PShuffleJoin.step over generator(42) inputs, sharded by batch range. Writes this
    rank's output rows to out_dir/out_{rank}.npy (the parent compares the union with the oracle)."""
    try:
        import numpy as np
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from dpu_olap_b200.generator import RandomArrayGenerator
        from dpu_olap_b200.ops import Context
        from dpu_olap_b200.sharded import P2PShuffleJoin
        ctx = Context(rank)
        per = nb_total // world
        first = rank * per
        g = RandomArrayGenerator(ctx, 42)
        x = g.batches_dev(nb_total, batch, take=(first, per))
        pk = g.index_column_dev(nb_total, batch, take=(first, per))
        y = g.batches_dev(nb_total, batch, take=(first, per))
        fk = g.foreign_key_dev(batch, nb_total, batch, take=(first, per))
        if dup:  # duplicate build keys + probe keys without a match: Arrow inner-join semantics
            pk = pk >> 1 << 1          # every even key twice, odd keys absent
        n = per * batch
        cap = n + n // 4 + 65536
        mult = 2 if dup else 1
        pj = P2PShuffleJoin(ctx, dist, rank, world, n, cap, probe_shares=shares)
        outs = [torch.empty(cap * mult, dtype=torch.int32, device="cuda") for _ in range(3)]
        rows_t = torch.empty(1, dtype=torch.int64, device="cuda")
        jws = torch.empty(ctx.join_seg_cap_ws_bytes(cap, cap, pj.nr_expected, pj.skip, pj.seg_bits) + 256,
                          dtype=torch.uint8, device="cuda")

        def local_join(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, skip_bits, abort, phase_bits):
            ctx.join_pairs_seg_cap_dev(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, out_capacity=cap * mult,
                                       skip_bits=skip_bits, ws=jws, outs=outs, out_rows=rows_t, abort=abort,
                                       phases=phase_bits)

        for _ in range(3):  # repeated steps reuse the receive buffers: the barrier protocol must hold
            outs[0].zero_()
            pj.step(fk, y, pk, x, local_join)
        torch.cuda.synchronize()
        m = ctx.join_rows(rows_t)
        res = np.stack([t[:m].cpu().numpy().view(np.uint32) for t in outs])
        np.save(os.path.join(out_dir, f"out_{rank}.npy"), res)
        np.save(os.path.join(out_dir, f"in_{rank}.npy"),
                np.stack([t.cpu().numpy().view(np.uint32) for t in (fk, y, pk, x)]))
        dist.barrier()
        dist.destroy_process_group()
        ctx.close()
    except Exception:  # noqa: BLE001 - the parent reads the file
        with open(os.path.join(out_dir, f"error_{rank}.txt"), "w") as f:
            f.write(traceback.format_exc())
        raise
