"""The reference's "Native" operators on Apache Arrow Acero — TEST INFRASTRUCTURE ONLY.

The reference's CPU path is Arrow C++ (apache-arrow-8.0.0, fetched by CPM, not vendored). The
only Arrow in this image is 24.0.0 inside the pyarrow wheel, whose Python Acero bindings drive the
same C++ engine. Each class below re-expresses one reference class with the same plan and options:

    FilterNative  host/filter/filter_native.cc:36-84   source -> filter(v < 1<<30) -> sink
    AggrNative    host/aggr/aggr_native.cc:39-93       source -> aggregate("sum", "v") -> sink
    TakeNative    host/take/take_native.cc:18-38       compute.take(values, indices, boundscheck=False) per batch
    JoinNative    host/join/join_native.cc:14-102      hashjoin INNER fk = pk, then drop pk

Used (a) to pin the C oracle on generator(42) inputs (tests/golden/make_arrow_golden.py) and
(b) as the CPU arm bench.py times next to the GPU (`--impl reference`, `cpu_baseline`).
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa
import pyarrow.acero as ac
import pyarrow.compute as pc

FILTER_THRESHOLD = 1 << 30


def _table(columns: dict[str, list[np.ndarray]]) -> pa.Table:
    """One record batch per input batch, zero-copy over the numpy buffers."""
    names = list(columns)
    nb = len(columns[names[0]])
    schema = pa.schema([pa.field(n, pa.uint32(), nullable=False) for n in names])
    batches = [pa.record_batch([pa.array(columns[n][b], type=pa.uint32()) for n in names], schema=schema)
               for b in range(nb)]
    return pa.Table.from_batches(batches, schema=schema)


class FilterNative:
    def __init__(self, batches: list[np.ndarray], threshold: int = FILTER_THRESHOLD, use_threads: bool = True):
        self.table = _table({"v": batches})
        self.threshold = threshold
        self.use_threads = use_threads

    def Prepare(self) -> None:
        expr = pc.less(pc.field("v"), pa.scalar(self.threshold, pa.uint32() if self.threshold < 2**32 else pa.uint64()))
        self.plan = ac.Declaration.from_sequence([
            ac.Declaration("table_source", ac.TableSourceNodeOptions(self.table)),
            ac.Declaration("filter", ac.FilterNodeOptions(expr)),
        ])

    def GetResult(self) -> pa.Table:
        return self.plan.to_table(use_threads=self.use_threads)

    def Run(self) -> int:
        return self.GetResult().num_rows


class AggrNative:
    def __init__(self, batches: list[np.ndarray], fn: str = "sum", use_threads: bool = True):
        self.table = _table({"v": batches})
        self.fn = fn
        self.use_threads = use_threads

    def Prepare(self) -> None:
        self.plan = ac.Declaration.from_sequence([
            ac.Declaration("table_source", ac.TableSourceNodeOptions(self.table)),
            ac.Declaration("aggregate", ac.AggregateNodeOptions([("v", self.fn, None, f"{self.fn}(v)")])),
        ])

    def Run(self) -> int:
        return int(self.plan.to_table(use_threads=self.use_threads).column(0)[0].as_py())


class TakeNative:
    def __init__(self, batches: list[np.ndarray], indices_batches: list[np.ndarray]):
        self.values = [pa.array(b, type=pa.uint32()) for b in batches]
        self.indices = [pa.array(b, type=pa.uint32()) for b in indices_batches]

    def Prepare(self) -> None:
        pass

    def Run(self) -> list[np.ndarray]:
        # the reference submits one Take per batch to the CPU thread pool (take_native.cc:22-33)
        from concurrent.futures import ThreadPoolExecutor
        def one(i):
            return pc.take(self.values[i], self.indices[i], boundscheck=False).to_numpy(zero_copy_only=False)
        with ThreadPoolExecutor(max_workers=pa.cpu_count()) as ex:
            return list(ex.map(one, range(len(self.values))))


class JoinNative:
    def __init__(self, left: dict[str, list[np.ndarray]], right: dict[str, list[np.ndarray]],
                 fk: str = "fk", pk: str = "pk", use_threads: bool = True):
        self.left, self.right = _table(left), _table(right)
        self.fk, self.pk = fk, pk
        self.use_threads = use_threads

    def Prepare(self) -> None:
        opts = ac.HashJoinNodeOptions("inner", left_keys=[self.fk], right_keys=[self.pk],
                                      output_suffix_for_left="_l", output_suffix_for_right="_r")
        self.plan = ac.Declaration("hashjoin", opts, inputs=[
            ac.Declaration("table_source", ac.TableSourceNodeOptions(self.left)),
            ac.Declaration("table_source", ac.TableSourceNodeOptions(self.right)),
        ])

    def Run(self) -> pa.Table:
        t = self.plan.to_table(use_threads=self.use_threads)
        return t.drop_columns([self.pk])  # join_native.cc:75
