"""Real multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

The fused NVLink shuffle (P2PShuffleJoin: count -> all-gather -> plan -> peer-store scatter -> barrier ->
segmented local join) is run by one process per GPU over NCCL and compared with oracle.join as a
sorted multiset, mirroring JoinTest.LargeTest (join_test.cc:82-121) at 32 x 65536 rows per side.
The one-GPU suites only reach this path with virtual ranks on one device."""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle

sys.path.insert(0, str(Path(__file__).resolve().parent))
import _mp_workers  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ngpus() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,dup,shares", [(2, False, 1), (2, True, 1), (2, False, 2), (2, True, 3), (4, False, None),
                                              (8, False, None)])
def test_p2p_shuffle_join_real_ranks(tmp_path, world, dup, shares):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    nb, batch = 32, 65536
    # shares: the probe side crosses NVLink in that many parts, each joined while the next is in flight
    mp.spawn(_mp_workers.p2p_join_worker, args=(world, _free_port(), nb, batch, str(tmp_path), dup, shares),
             nprocs=world, join=True)
    errs = sorted(tmp_path.glob("error_*.txt"))
    assert not errs, errs[0].read_text()
    ins = [np.load(tmp_path / f"in_{r}.npy") for r in range(world)]
    fk, y, pk, x = (np.concatenate([i[c] for i in ins]) for c in range(4))
    got = np.concatenate([np.load(tmp_path / f"out_{r}.npy") for r in range(world)], axis=1)
    exp = oracle.join(fk, y, pk, x)
    assert got.shape[1] == exp[0].size
    if not dup:
        assert got.shape[1] == nb * batch  # JoinTest.LargeTest: every probe row matches once
    for a, b in zip(oracle.sort_rows(*got), oracle.sort_rows(*exp)):
        assert np.array_equal(a, b)
