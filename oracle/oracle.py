"""ctypes/numpy front end of oracle/olap_oracle.c (the C restatement) — test infrastructure only.

Every function cites the reference lines it follows in olap_oracle.c. The sorted-multiset
helpers mirror how the reference's own tests compare join results (join_test.cc:27-38,76-77).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"

__all__ = ["build", "RandomArrayGenerator", "gen_u32", "iota_u32", "filter_lt", "sum_u32", "take",
           "wang_hash", "bucket", "partition_ids", "join", "sort_rows", "triple_checksum",
           "make_random_batches", "make_fk_batches", "make_index_batches",
           "join_aggr", "filter_lt_typed", "filter_lt_nullable", "aggr_nullable", "take_nullable", "pack_bits", "unpack_bits"]

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "olap_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "liboracle.so"], check=True,
                       capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        u32p, i64, u64, u32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32
        L.orc_minstd_seed.argtypes = [C.c_void_p, u32]
        L.orc_next_seed.argtypes = [C.c_void_p]
        L.orc_next_seed.restype = C.c_int32
        L.orc_gen_u32.argtypes = [i64, u32, u32, i64, u32p]
        L.orc_iota_u32.argtypes = [u64, i64, u32p]
        L.orc_filter_lt_u32.argtypes = [u32p, i64, u32, u32p]
        L.orc_filter_lt_u32.restype = i64
        L.orc_sum_u32.argtypes = [u32p, i64]
        L.orc_sum_u32.restype = u64
        L.orc_take_u32.argtypes = [u32p, u32p, i64, u32p]
        L.orc_wang_hash_u32.argtypes = [u32]
        L.orc_wang_hash_u32.restype = u32
        L.orc_bucket.argtypes = [u32, u32, C.c_int]
        L.orc_bucket.restype = u32
        L.orc_partition_ids.argtypes = [u32p, i64, u32, C.c_int, u32p]
        L.orc_join_u32.argtypes = [u32p, u32p, i64, u32p, u32p, i64, u32p, u32p, u32p, i64]
        L.orc_join_u32.restype = i64
        L.orc_triple_checksum.argtypes = [u32p, u32p, u32p, i64]
        L.orc_triple_checksum.restype = u64
        _lib = L
    return _lib


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


class RandomArrayGenerator:
    """arrow::random::RandomArrayGenerator's seed stream (arrow/testing/random.h) + uint32 arrays."""

    def __init__(self, seed: int = 42):
        self._state = C.c_uint64(0)
        lib().orc_minstd_seed(C.byref(self._state), seed)

    def seed(self) -> int:
        return int(lib().orc_next_seed(C.byref(self._state)))

    def data_seed(self) -> int:
        """Seed fed to pcg32_fast for the DATA of the next array: seed() + 1 as int32
        (GenerateBitmap consumed seed_ first, random.cc:111-125,190-196)."""
        s = self.seed() + 1
        if s > 0x7FFFFFFF:
            s -= 1 << 32
        return s

    def uint32(self, n: int, lo: int = 0, hi: int = 0xFFFFFFFF) -> np.ndarray:
        return gen_u32(self.data_seed(), n, lo, hi)


def gen_u32(data_seed: int, n: int, lo: int = 0, hi: int = 0xFFFFFFFF) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().orc_gen_u32(int(data_seed), lo, hi, n, out.ctypes.data)
    return out


def iota_u32(start: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().orc_iota_u32(start, n, out.ctypes.data)
    return out


# generator::MakeRandomRecordBatches / MakeForeignKeyColumn / MakeIndexColumn (generator.cc:22-71)
def make_random_batches(g: RandomArrayGenerator, num_batches: int, batch_size: int, lo: int = 0,
                        hi: int = 0xFFFFFFFF) -> list[np.ndarray]:
    return [g.uint32(batch_size, lo, hi) for _ in range(num_batches)]


def make_fk_batches(g: RandomArrayGenerator, pk_batch_size: int, num_batches: int,
                    batch_size: int) -> list[np.ndarray]:
    return [g.uint32(batch_size, (i * pk_batch_size) & 0xFFFFFFFF, ((i + 1) * pk_batch_size - 1) & 0xFFFFFFFF)
            for i in range(num_batches)]


def make_index_batches(num_batches: int, batch_size: int) -> list[np.ndarray]:
    return [iota_u32(i * batch_size, batch_size) for i in range(num_batches)]


def filter_lt(col, thr: int = 1 << 30) -> np.ndarray:
    a = _u32(col)
    out = np.empty(a.size, dtype=np.uint32)
    m = lib().orc_filter_lt_u32(a.ctypes.data, a.size, thr, out.ctypes.data)
    return out[:m].copy()


def sum_u32(col) -> int:
    a = _u32(col)
    return int(lib().orc_sum_u32(a.ctypes.data, a.size))


def take(values, indices) -> np.ndarray:
    v, i = _u32(values), _u32(indices)
    out = np.empty(i.size, dtype=np.uint32)
    lib().orc_take_u32(v.ctypes.data, i.ctypes.data, i.size, out.ctypes.data)
    return out


def wang_hash(key: int) -> int:
    return int(lib().orc_wang_hash_u32(key & 0xFFFFFFFF))


def bucket(key: int, nparts: int, skip_bits: int = 0) -> int:
    return int(lib().orc_bucket(key & 0xFFFFFFFF, nparts, skip_bits))


def partition_ids(keys, nparts: int, skip_bits: int = 0) -> np.ndarray:
    k = _u32(keys)
    out = np.empty(k.size, dtype=np.uint32)
    lib().orc_partition_ids(k.ctypes.data, k.size, nparts, skip_bits, out.ctypes.data)
    return out


def join(fk, y, pk, x):
    """Inner join; returns (fk, y, x) arrays in probe order (compare as sorted multisets)."""
    fk, y, pk, x = _u32(fk), _u32(y), _u32(pk), _u32(x)
    cap = max(fk.size, 1)
    while True:
        o = [np.empty(cap, dtype=np.uint32) for _ in range(3)]
        m = lib().orc_join_u32(fk.ctypes.data, y.ctypes.data, fk.size, pk.ctypes.data, x.ctypes.data,
                               pk.size, o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data, cap)
        if m < 0:
            raise MemoryError("oracle join")
        if m <= cap:
            return tuple(a[:m].copy() for a in o)
        cap = m


def join_aggr(fk, y, pk, x, y_threshold=None) -> dict:
    """SELECT COUNT(*), SUM(L.y), SUM(R.x) FROM L JOIN R ON fk = pk [WHERE L.y < y_threshold]: the
    reference's Native operators composed (filter_native.cc:52-66 on the probe side, join_native.cc:31-40,
    aggr_native.cc:68-73 over the result columns); sums modulo 2^64."""
    fk, y = _u32(fk), _u32(y)
    if y_threshold is not None:
        keep = y < np.uint32(y_threshold) if y_threshold <= 0xFFFFFFFF else np.ones(y.size, bool)
        fk, y = fk[keep], y[keep]
    o_fk, o_y, o_x = join(fk, y, pk, x)
    return {"rows": int(o_fk.size), "sum_y": int(o_y.astype(np.uint64).sum(dtype=np.uint64)),
            "sum_x": int(o_x.astype(np.uint64).sum(dtype=np.uint64))}


def sort_rows(*cols):
    """Rows sorted lexicographically by (col0, col1, ...) — the multiset normal form."""
    cols = [_u32(c) for c in cols]
    if cols[0].size == 0:
        return tuple(cols)
    order = np.lexsort(tuple(reversed(cols)))
    return tuple(c[order] for c in cols)


def triple_checksum(a, b, c) -> int:
    a, b, c = _u32(a), _u32(b), _u32(c)
    return int(lib().orc_triple_checksum(a.ctypes.data, b.ctypes.data, c.ctypes.data, a.size))


# ---- nullable columns -------------------------------------------------------
# The reference's DPU kernels know no nulls (nullptr bitmap, filter_dpu.cc:91); its ORACLE does:
# the Native classes call Arrow's filter / sum / take (filter_native.cc:52-66, aggr_native.cc:68-73,
# take_native.cc:27), whose null handling is restated here in numpy over (values, valid) pairs,
# valid[i] = True where row i is non-null. tests/test_oracle_nullable.py pins these against
# pyarrow.compute (Arrow 24), the same library the reference's plans run on.
def filter_lt_nullable(values, valid, thr: int = 1 << 30) -> np.ndarray:
    """Acero filter(less(v, thr)): a null predicate drops the row, so the result has no nulls."""
    v = _u32(values)
    keep = np.asarray(valid, dtype=bool) & (v < np.uint32(thr) if thr <= 0xFFFFFFFF else np.ones(v.size, bool))
    return v[keep].copy()


def filter_lt_typed(values, valid, thr) -> np.ndarray:
    """The same filter over a uint32 / int32 / float32 or uint64 / int64 / float64 column in ITS type's
    order (Arrow's `less`: signed for the int types, IEEE for the float types — a NaN row, like a null
    one, is never selected)."""
    v = np.ascontiguousarray(values)
    assert v.dtype in (np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64)
    with np.errstate(invalid="ignore"):
        keep = np.asarray(valid, dtype=bool) & (v < np.asarray(thr, dtype=v.dtype))
    return v[keep].copy()


def aggr_nullable(values, valid) -> dict:
    """cp::Sum / Count / MinMax with default options: nulls are skipped; with no valid row every
    aggregate but the count is null (None)."""
    v = np.ascontiguousarray(values)
    if v.dtype in (np.uint64, np.int64):
        # 64-bit columns: Arrow's sum has the column's type and wraps (unchecked), min / max in the
        # type's order
        v = v[np.asarray(valid, dtype=bool)]
        if v.size == 0:
            return {"sum": None, "count": 0, "min": None, "max": None}
        total = int(v.view(np.uint64).sum(dtype=np.uint64))  # mod 2^64
        if v.dtype == np.int64 and total >> 63:
            total -= 1 << 64
        return {"sum": total, "count": int(v.size), "min": int(v.min()), "max": int(v.max())}
    if v.dtype != np.int32:   # int32 columns sum into int64 and compare signed; everything else is uint32
        v = _u32(v)
    v = v[np.asarray(valid, dtype=bool)]
    if v.size == 0:
        return {"sum": None, "count": 0, "min": None, "max": None}
    wide = np.int64 if v.dtype == np.int32 else np.uint64
    return {"sum": int(v.astype(wide).sum(dtype=wide)), "count": int(v.size),
            "min": int(v.min()), "max": int(v.max())}


def take_nullable(values, values_valid, indices, indices_valid):
    """cp::Take(values, indices): slot j is null when index j is null or values[indices[j]] is null.
    Returns (out, out_valid); null slots of out are 0 (the CUDA kernel's convention)."""
    v, i = _u32(values), _u32(indices)
    vv, iv = np.asarray(values_valid, dtype=bool), np.asarray(indices_valid, dtype=bool)
    safe = np.where(iv, i, 0).astype(np.int64)
    ok = iv & (vv[safe] if v.size else np.zeros(i.size, bool))
    out = np.where(ok, v[safe] if v.size else 0, 0).astype(np.uint32)
    return out, ok


def pack_bits(valid) -> np.ndarray:
    """Arrow validity bitmap (LSB first) of a boolean mask, padded to a multiple of 4 bytes + 4."""
    b = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    out = np.zeros((b.size + 3) // 4 * 4 + 4, dtype=np.uint8)
    out[:b.size] = b
    return out


def unpack_bits(bitmap, n: int, offset: int = 0) -> np.ndarray:
    return np.unpackbits(np.asarray(bitmap, dtype=np.uint8), bitorder="little")[offset:offset + n].astype(bool)
