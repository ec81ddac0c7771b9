"""Synthetic inputs of the benchmark harness — mirror of dpu_olap's host/generator.

The reference draws every column from ``arrow::random::RandomArrayGenerator(42)``
(host/generator/generator.cc:22-71, host/*/…_benchmark.cc fixtures). An array's values depend
only on its own 32-bit seed, taken in order from the generator's seed stream
(arrow/testing/random.h: ``std::default_random_engine seed_rng(seed)``,
``std::uniform_int_distribution<int32_t>(1, INT32_MAX)``); the data itself is produced ON THE GPU
by ``b2_gen_u32_dev`` (csrc/gen.cu), bit-identical to the host generator, so SF=2048 columns
(64 GiB) never pass through host memory.

This module only walks the seed stream (a few hundred thousand LCG steps at most) and lays the
fixtures out; it computes no column data.
"""
from __future__ import annotations

import numpy as np

_M = 2147483647  # minstd_rand0 modulus (std::default_random_engine)
_A = 16807


class SeedStream:
    """RandomArrayGenerator::seed() — libstdc++'s uniform_int_distribution<int32_t>(1, INT32_MAX)
    over minstd_rand0 (URNG range 2^31-3 is one short of the target range, so libstdc++ takes its
    "upscaling" branch: a high part drawn from {0,1} by rejection and a low raw draw)."""

    def __init__(self, seed: int = 42):
        s = seed % _M
        self._x = s if s else 1

    def _next(self) -> int:
        self._x = (self._x * _A) % _M
        return self._x

    def _uniform(self, urange: int) -> int:
        urngrange = _M - 2  # (m - 1) - 1
        if urngrange > urange:
            uerange = urange + 1
            scaling = urngrange // uerange
            past = uerange * scaling
            while True:
                ret = self._next() - 1
                if ret < past:
                    return ret // scaling
        if urngrange < urange:
            uerngrange = urngrange + 1
            while True:
                tmp = uerngrange * self._uniform(urange // uerngrange)
                ret = tmp + (self._next() - 1)
                if tmp <= ret <= urange:
                    return ret
        return self._next() - 1

    def seed(self) -> int:
        return self._uniform(0x7FFFFFFF - 1) + 1

    def data_seed(self) -> int:
        """Seed fed to pcg32_fast for an array's DATA: seed()+1 as int32 — the validity bitmap is
        drawn first and consumes ``seed_++`` (host/generator/random.cc:111-125,190-196)."""
        s = self.seed() + 1
        return s - (1 << 32) if s > 0x7FFFFFFF else s

    def data_seeds(self, n: int) -> np.ndarray:
        """Next n data seeds as the uint64 values pcg32_fast receives (int32 sign-extended)."""
        return np.array([self.data_seed() & 0xFFFFFFFFFFFFFFFF for _ in range(n)], dtype=np.uint64)

    def skip(self, n: int) -> None:
        for _ in range(n):
            self.seed()


class RandomArrayGenerator(SeedStream):
    """Device-side counterpart of arrow::random::RandomArrayGenerator for uint32 columns."""

    def __init__(self, ctx, seed: int = 42):
        super().__init__(seed)
        self.ctx = ctx

    def batches_dev(self, num_batches: int, batch_size: int, lo=None, hi=None, take=None, out=None):
        """generator::MakeRandomRecordBatches for a one-column uint32 schema, packed on the device.

        ``take=(first, count)`` materialises only that range of batches (row-range sharding across
        GPUs) while still consuming all num_batches seeds, so every rank sees the same stream."""
        seeds = self.data_seeds(num_batches)
        first, count = (0, num_batches) if take is None else take
        sel = slice(first, first + count)
        lo_a = hi_a = None
        if lo is not None:
            lo_a = np.broadcast_to(np.asarray(lo, dtype=np.uint32), (num_batches,))[sel]
            hi_a = np.broadcast_to(np.asarray(hi, dtype=np.uint32), (num_batches,))[sel]
        return self.ctx.gen_dev(seeds[sel], count, batch_size, lo_a, hi_a, out=out)

    def foreign_key_dev(self, pk_batch_size: int, num_batches: int, batch_size: int, take=None):
        """generator::MakeForeignKeyColumn (generator.cc:46-57): batch i uniform in
        [i*pk_batch_size, (i+1)*pk_batch_size - 1] (uint32 arithmetic, wraps like the reference)."""
        i = np.arange(num_batches, dtype=np.uint64)
        lo = ((i * pk_batch_size) & 0xFFFFFFFF).astype(np.uint32)
        hi = (((i + 1) * pk_batch_size - 1) & 0xFFFFFFFF).astype(np.uint32)
        return self.batches_dev(num_batches, batch_size, lo, hi, take=take)

    def index_column_dev(self, num_batches: int, batch_size: int, take=None):
        """generator::MakeIndexColumn (generator.cc:59-71): 0,1,2,... across batches."""
        first, count = (0, num_batches) if take is None else take
        return self.ctx.iota_dev(first * batch_size, count * batch_size)
