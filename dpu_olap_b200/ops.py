"""Host-side mirror of dpu_olap's operator API for the GPU path.

Same duck-typed surface as the reference's *Dpu classes — constructor, ``Prepare()``, ``Run()``
(+ ``GetResult()`` for the filter), ``Timers()`` — with the DpuSet replaced by a :class:`Context`
bound to one B200:

    reference (C++)                                   here
    FilterDpu(system, batches)      filter_dpu.h:14   FilterGpu(ctx, batches)
    SumDpu(system, batches)         aggr_dpu.h:14     SumGpu(ctx, batches)
    TakeDpu(system, b, idx)         take_dpu.h:14     TakeGpu(ctx, batches, indices_batches)
    JoinDpu(system, ls, rs, lb, rb) join_dpu.h:14     JoinGpu(ctx, left_batches, right_batches)
    PartitionDpu(system, s, b, n, k) partition_dpu.h:15  PartitionGpu(ctx, batches, nr_partitions, key)

Batches are "record batches" in the loosest useful sense: a ``pyarrow.RecordBatch``, a dict
``{column name: uint32 array}`` or (single-column operators) a bare uint32 ``numpy``/``pyarrow``
array. Only the raw data buffers cross the C ABI (include/b200olap.h); nothing here computes —
if the CUDA library is missing these classes cannot be constructed.

The ``*_dev`` functions of :class:`Context` are the device-resident entry points (torch CUDA
tensors in, torch CUDA tensors out) used by bench.py and the parity tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

import numpy as np

from . import _lib
from ._lib import B2Error, JoinPhases, Timings, check

try:  # pyarrow is optional for the package, present in the image
    import pyarrow as pa
except Exception:  # pragma: no cover
    pa = None


class Timers(dict):
    """Named timers in milliseconds, reference names (filter_dpu.cc:107-110)."""

    @classmethod
    def from_timings(cls, *ts: Timings) -> "Timers":
        out = cls({"copy-to-dpu": 0.0, "dpu-work": 0.0, "copy-from-dpu": 0.0, "total": 0.0})
        for t in ts:
            out["copy-to-dpu"] += t.copy_to_dev_ms
            out["dpu-work"] += t.dev_work_ms
            out["copy-from-dpu"] += t.copy_from_dev_ms
            out["total"] += t.total_ms
        return out

    def add_join_phases(self, p: dict) -> None:
        """JoinDpu's timer names (join_dpu.cc:146-148). The probe kernel builds each partition's table
        and probes it in one launch, so "probe" covers the reference's "build" + "probe"."""
        self["partitionKernel"] = p["partition_build_ms"] + p["partition_probe_ms"]
        self["partitionKernel-build-side"] = p["partition_build_ms"]
        self["partitionKernel-probe-side"] = p["partition_probe_ms"]
        self["probe"] = p["probe_ms"]
        self["take"] = p["take_ms"]

    def get(self, *a):  # reference: timers->get() returns the map
        return self if not a else dict.get(self, *a)


# ---------------------------------------------------------------------------------------------
# batch plumbing
# ---------------------------------------------------------------------------------------------
def _as_u32(col: Any) -> np.ndarray:
    """Zero-copy uint32 numpy view of one column of one batch."""
    if pa is not None and isinstance(col, (pa.Array, pa.ChunkedArray)):
        if isinstance(col, pa.ChunkedArray):
            col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
        if col.null_count:
            raise ValueError("nullable columns are not supported (the reference assumes non-null)")
        if col.type != pa.uint32():
            raise TypeError(f"expected uint32 column, got {col.type}")
        return col.to_numpy(zero_copy_only=True)
    a = np.asarray(col)
    if a.dtype != np.uint32:
        raise TypeError(f"expected uint32 column, got {a.dtype}")
    return np.ascontiguousarray(a)


def _column(batch: Any, name_or_index) -> np.ndarray:
    if pa is not None and isinstance(batch, pa.RecordBatch):
        if isinstance(name_or_index, str):
            return _as_u32(batch.column(batch.schema.get_field_index(name_or_index)))
        return _as_u32(batch.column(name_or_index))
    if isinstance(batch, dict):
        if isinstance(name_or_index, str):
            return _as_u32(batch[name_or_index])
        return _as_u32(list(batch.values())[name_or_index])
    if name_or_index not in (0, None):
        raise KeyError("a bare array batch has a single column")
    return _as_u32(batch)


def _raw_column(batch: Any, name: str) -> np.ndarray:
    """One non-null column of one batch as a contiguous numpy array of ITS OWN fixed-width type
    (4 or 8 bytes per element): what b2_join_table_host moves as raw words."""
    if pa is not None and isinstance(batch, pa.RecordBatch):
        col = batch.column(batch.schema.get_field_index(name))
        if col.null_count:
            raise ValueError("the typed-table join takes non-null columns")
        a = col.to_numpy(zero_copy_only=True)
    else:
        col = batch[name]
        if pa is not None and isinstance(col, (pa.Array, pa.ChunkedArray)):
            if col.null_count:
                raise ValueError("the typed-table join takes non-null columns")
            col = (col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col).to_numpy(zero_copy_only=True)
        a = np.ascontiguousarray(np.asarray(col))
    if a.dtype.kind not in "uif" or a.dtype.itemsize not in (4, 8):
        raise TypeError(f"column {name!r}: expected a 32- or 64-bit numeric type, got {a.dtype}")
    return a


def _column_dtype(batch: Any, name: str) -> np.dtype:
    if pa is not None and isinstance(batch, pa.RecordBatch):
        return np.dtype(batch.schema.field(name).type.to_pandas_dtype())
    col = batch[name]
    if pa is not None and isinstance(col, (pa.Array, pa.ChunkedArray)):
        return np.dtype(col.type.to_pandas_dtype())  # not np.asarray: nulls would turn integers into floats
    if isinstance(col, np.ma.MaskedArray):
        return np.ma.getdata(col).dtype
    return np.asarray(col).dtype


# 32-bit column types the filter compares natively (b2_dtype32, include/b200olap.h)
_DTYPES32 = {np.dtype(np.uint32): 0, np.dtype(np.int32): 1, np.dtype(np.float32): 2}
_DTYPES64 = {np.dtype(np.uint64): 3, np.dtype(np.int64): 4}  # b2_dtype64 (aggregates, take, filter)
_DTYPES64_FILTER = {**_DTYPES64, np.dtype(np.float64): 5}       # the filter also compares float64


def _threshold_bits(threshold, dtype: np.dtype) -> int:
    """Bit pattern of the threshold in the column's type."""
    dtype = np.dtype(dtype)
    return int(np.array([threshold], dtype=dtype).view(np.uint64 if dtype.itemsize == 8 else np.uint32)[0])


class _NullableCol:
    """One column of one batch that may carry an Arrow validity bitmap:
    ``values`` is the uint32 data (null slots hold whatever the buffer holds), ``valid`` a uint8
    array holding the bitmap (bit ``offset + r`` = row r, LSB first) or None when there are no nulls."""

    __slots__ = ("values", "valid", "offset", "dtype", "_keep")

    def __init__(self, values: np.ndarray, valid=None, offset: int = 0, keep=None):
        self.dtype = values.dtype                       # the column's own type
        # what crosses the ABI: raw 32-bit words (64-bit columns, aggregates only: raw 64-bit words)
        self.values = values.view(np.uint64 if values.dtype.itemsize == 8 else np.uint32)
        self.valid, self.offset, self._keep = valid, int(offset), keep


def _as_nullable(col: Any, typed: bool = False, wide: bool = False, f64: bool = False) -> _NullableCol:
    """typed=True also admits int32 / float32 columns (filter, aggregates); wide=True also uint64 /
    int64 columns (aggregates, take, filter); f64=True also float64 (filter)."""
    ok = tuple(_DTYPES32) if typed else (np.dtype(np.uint32),)
    if wide:
        ok = ok + tuple(_DTYPES64)
    if f64:
        ok = ok + (np.dtype(np.float64),)
    if pa is not None and isinstance(col, (pa.Array, pa.ChunkedArray)):
        if isinstance(col, pa.ChunkedArray):
            col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
        try:
            dt = np.dtype(col.type.to_pandas_dtype())
        except NotImplementedError:
            dt = None
        if dt not in ok:
            raise TypeError(f"expected a {' / '.join(str(d) for d in ok)} column, got {col.type}")
        vbuf, dbuf = col.buffers()[0], col.buffers()[1]
        values = (np.frombuffer(dbuf, dtype=dt)[col.offset:col.offset + len(col)] if dbuf is not None
                  else np.empty(0, dtype=dt))
        if col.null_count == 0:
            return _NullableCol(values, keep=col)
        return _NullableCol(values, np.frombuffer(vbuf, dtype=np.uint8), col.offset, keep=col)
    if isinstance(col, np.ma.MaskedArray):
        data = np.ascontiguousarray(np.ma.getdata(col))
        if data.dtype not in ok:
            raise TypeError(f"expected a {' / '.join(str(d) for d in ok)} column, got {data.dtype}")
        mask = np.ma.getmaskarray(col)
        if not mask.any():
            return _NullableCol(data)
        return _NullableCol(data, np.packbits(~mask, bitorder="little"), 0)
    a = np.asarray(col)
    if a.dtype not in ok:
        raise TypeError(f"expected a {' / '.join(str(d) for d in ok)} column, got {a.dtype}")
    return _NullableCol(np.ascontiguousarray(a))


def _nullable_column(batch: Any, name_or_index, typed: bool = False, wide: bool = False,
                     f64: bool = False) -> _NullableCol:
    if pa is not None and isinstance(batch, pa.RecordBatch):
        i = batch.schema.get_field_index(name_or_index) if isinstance(name_or_index, str) else name_or_index
        return _as_nullable(batch.column(i), typed, wide, f64)
    if isinstance(batch, dict):
        return _as_nullable(batch[name_or_index] if isinstance(name_or_index, str)
                            else list(batch.values())[name_or_index], typed, wide, f64)
    return _as_nullable(batch, typed, wide, f64)


class _ValidTable:
    """(const uint8_t* const* bitmaps, const int64_t* bit offsets) of a list of _NullableCol."""

    def __init__(self, cols: Sequence[_NullableCol]):
        n = len(cols)
        self.any = any(c.valid is not None for c in cols)
        self.ptrs = (C.c_void_p * max(n, 1))(*[c.valid.ctypes.data if c.valid is not None else None
                                               for c in cols])
        self.offs = (C.c_int64 * max(n, 1))(*[c.offset for c in cols])
        self._keep = cols


def _to_arrow(values: np.ndarray, valid_bytes: np.ndarray):
    """uint32 pyarrow array over (data, validity bitmap with bit offset 0); no copy."""
    return pa.Array.from_buffers(pa.uint32(), values.size, [pa.py_buffer(valid_bytes), pa.py_buffer(values)])


def _column_names(batch: Any) -> list[str]:
    if pa is not None and isinstance(batch, pa.RecordBatch):
        return list(batch.schema.names)
    if isinstance(batch, dict):
        return list(batch.keys())
    return ["v"]


class _PtrTable:
    """(const uint32_t* const*, const int64_t*) view of a list of arrays; keeps them alive."""

    def __init__(self, arrays: Sequence[np.ndarray]):
        self.arrays = list(arrays)
        n = len(self.arrays)
        self.ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in self.arrays])
        self.lens = (C.c_int64 * max(n, 1))(*[int(a.size) for a in self.arrays])
        self.n = n


def _dptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
class Context:
    """One B200. Replaces ``dpu::DpuSet::allocate(nr_dpus)`` (dpuext.hpp:710)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self._lib = _lib.lib()  # raises if libb200olap.so is absent — no fallback
        self._owned = _borrowed is None
        if _borrowed is None:
            h = C.c_void_p()
            check(self._lib.b2_ctx_create(int(device), C.byref(h)), "b2_ctx_create")
            self._h = h
            self.device = int(device)
        else:  # a member of a DeviceSet: the set owns it
            self._h = C.c_void_p(_borrowed)
            self.device = int(self._lib.b2_ctx_device(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.b2_ctx_destroy(self._h)
            self._h = None

    def _call(self, name: str, *args) -> None:
        """b2_<name>(ctx, ...): the host entry point of an operator on this one GPU."""
        self._ck(getattr(self._lib, "b2_" + name)(self._h, *args), "b2_" + name)

    def join_trace(self, on: bool = True) -> None:
        """b2_join_trace: later joins of this context record CUDA events at their phase boundaries."""
        self._ck(self._lib.b2_join_trace(self._h, 1 if on else 0), "b2_join_trace")

    def join_last_phases(self) -> dict:
        """b2_join_last_phases: phase timers of the last traced join (waits for its events)."""
        p = JoinPhases()
        self._ck(self._lib.b2_join_last_phases(self._h, C.byref(p)), "b2_join_last_phases")
        return p.as_dict()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, status: int, where: str) -> None:
        check(status, where, self._h)

    @property
    def launches(self) -> int:
        return int(self._lib.b2_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(self._lib.b2_ctx_sm_count(self._h))

    def set_tunable(self, which: int, value: int) -> None:
        """Kernel-selection knob of this ctx (enum b2_tunable; every setting computes the same result)."""
        self._ck(self._lib.b2_ctx_set_tunable(self._h, int(which), int(value)), "b2_ctx_set_tunable")

    def get_tunable(self, which: int) -> int:
        v = C.c_int(0)
        self._ck(self._lib.b2_ctx_get_tunable(self._h, int(which), C.byref(v)), "b2_ctx_get_tunable")
        return int(v.value)

    # ---- device-resident entry points (torch CUDA tensors) --------------------------------
    @staticmethod
    def _stream() -> int:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    def gen_dev(self, data_seeds, nbatches: int, batch_len: int, lo=None, hi=None, out=None):
        """RandomArrayGenerator data for nbatches arrays (seeds = the pcg32_fast seeds)."""
        import torch
        if out is None:
            out = torch.empty(nbatches * batch_len, dtype=torch.int32, device=f"cuda:{self.device}")
        seeds = np.ascontiguousarray(np.asarray(data_seeds, dtype=np.uint64))
        assert seeds.size == nbatches
        lo_p = hi_p = None
        if lo is not None:
            lo_a = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.uint32), (nbatches,)))
            hi_a = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.uint32), (nbatches,)))
            lo_p = lo_a.ctypes.data_as(C.POINTER(C.c_uint32))
            hi_p = hi_a.ctypes.data_as(C.POINTER(C.c_uint32))
        self._ck(self._lib.b2_gen_u32_dev(self._h, seeds.ctypes.data_as(C.POINTER(C.c_uint64)), lo_p,
                                          hi_p, nbatches, batch_len, _dptr(out), self._stream()),
                 "b2_gen_u32_dev")
        return out

    def iota_dev(self, start: int, n: int, out=None):
        import torch
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=f"cuda:{self.device}")
        self._ck(self._lib.b2_iota_u32_dev(self._h, start, n, _dptr(out), self._stream()), "b2_iota_u32_dev")
        return out

    def sum_dev(self, col, out=None):
        import torch
        if out is None:
            out = torch.empty(1, dtype=torch.int64, device=col.device)
        self._ck(self._lib.b2_sum_u32_dev(self._h, _dptr(col), col.numel(), _dptr(out), self._stream()),
                 "b2_sum_u32_dev")
        return out

    def sum_lt_dev(self, col, threshold: int):
        """Fused filter(v < threshold) -> sum. Returns (sum, count) as uint64-in-int64 device tensors."""
        import torch
        out = torch.empty(1, dtype=torch.int64, device=col.device)
        cnt = torch.empty(1, dtype=torch.int64, device=col.device)
        self._ck(self._lib.b2_sum_lt_u32_dev(self._h, _dptr(col), col.numel(), int(threshold), _dptr(out),
                                             _dptr(cnt), self._stream()), "b2_sum_lt_u32_dev")
        return out, cnt

    def filter_ws_bytes(self, nbatches: int, batch_len: int) -> int:
        return int(self._lib.b2_filter_ws_bytes(nbatches, batch_len))

    def filter_dev(self, col, nbatches: int, batch_len: int, threshold: int, out=None, batch_end=None,
                   total=None, ws=None, carry_in=None):
        """Returns (out, batch_end, total) device tensors; nothing is synchronised."""
        import torch
        dev = col.device
        if out is None:
            out = torch.empty(nbatches * batch_len, dtype=torch.int32, device=dev)
        if batch_end is None:
            batch_end = torch.empty(max(nbatches, 1), dtype=torch.int64, device=dev)
        if total is None:
            total = torch.empty(1, dtype=torch.int64, device=dev)
        need = self.filter_ws_bytes(nbatches, batch_len)
        if ws is None:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._ck(self._lib.b2_filter_lt_u32_dev(self._h, _dptr(col), nbatches, batch_len, threshold,
                                                _dptr(out), _dptr(batch_end), _dptr(total),
                                                _dptr(carry_in), _dptr(ws), ws.numel(), self._stream()),
                 "b2_filter_lt_u32_dev")
        return out, batch_end, total

    def filter_nullable_dev(self, col, valid, nbatches: int, batch_len: int, threshold: int, out=None,
                            batch_end=None, total=None, ws=None):
        """Nullable filter: `valid` is the packed validity bitmap (uint8 device tensor, padded to 4
        bytes) or None. Returns (out, batch_end, total) device tensors."""
        import torch
        dev = col.device
        if out is None:
            out = torch.empty(max(nbatches * batch_len, 1), dtype=torch.int32, device=dev)
        if batch_end is None:
            batch_end = torch.empty(max(nbatches, 1), dtype=torch.int64, device=dev)
        if total is None:
            total = torch.empty(1, dtype=torch.int64, device=dev)
        if ws is None:
            ws = torch.empty(self.filter_ws_bytes(nbatches, batch_len), dtype=torch.uint8, device=dev)
        self._ck(self._lib.b2_filter_lt_u32_nullable_dev(self._h, _dptr(col), _dptr(valid), nbatches, batch_len,
                                                         threshold, _dptr(out), _dptr(batch_end), _dptr(total),
                                                         0, _dptr(ws), ws.numel(), self._stream()),
                 "b2_filter_lt_u32_nullable_dev")
        return out, batch_end, total

    def filter_typed_dev(self, col, dtype, threshold, nbatches: int, batch_len: int, valid=None):
        """Filter `v < threshold` over an int32 / float32 / uint32 column held as a 32-bit device tensor
        (b2_filter_lt_32_dev). Returns (out (same torch dtype as col), batch_end, total)."""
        import torch
        dt = np.dtype(dtype)
        dev = col.device
        out = torch.empty(max(nbatches * batch_len, 1), dtype=col.dtype, device=dev)
        batch_end = torch.empty(max(nbatches, 1), dtype=torch.int64, device=dev)
        total = torch.empty(1, dtype=torch.int64, device=dev)
        ws = torch.empty(self.filter_ws_bytes(nbatches, batch_len), dtype=torch.uint8, device=dev)
        self._ck(self._lib.b2_filter_lt_32_dev(self._h, _dptr(col), _DTYPES32[dt], _threshold_bits(threshold, dt),
                                               _dptr(valid), nbatches, batch_len, _dptr(out), _dptr(batch_end),
                                               _dptr(total), 0, _dptr(ws), ws.numel(), self._stream()),
                 "b2_filter_lt_32_dev")
        return out, batch_end, total

    def filter64_dev(self, col, dtype, threshold, valid=None, batch_off=None, nbatches: int = 1,
                     batch_len: int | None = None, out=None, batch_end=None, total=None, ws=None):
        """b2_filter_lt_64_dev over a device column of 64-bit values (an int64 torch tensor holding the bit
        patterns). Batches: batch_off (int64 device tensor, nbatches + 1) or nbatches x batch_len rows."""
        import torch
        dt = np.dtype(dtype)
        n = col.numel()
        if batch_off is None and batch_len is None:
            nbatches, batch_len = 1, n
        if out is None:
            out = torch.empty(max(n, 1), dtype=torch.int64, device=col.device)
        if batch_end is None:
            batch_end = torch.empty(max(nbatches, 1), dtype=torch.int64, device=col.device)
        if total is None:
            total = torch.empty(1, dtype=torch.int64, device=col.device)
        if ws is None:
            ws = torch.empty(int(self._lib.b2_filter_64_ws_bytes(n)) + 256, dtype=torch.uint8, device=col.device)
        ptr, nbytes = self._aligned(ws)
        self._ck(self._lib.b2_filter_lt_64_dev(self._h, _dptr(col), _DTYPES64_FILTER[dt], _threshold_bits(threshold, dt),
                                               _dptr(valid), n, _dptr(batch_off), nbatches,
                                               0 if batch_len is None else batch_len, _dptr(out), _dptr(batch_end),
                                               _dptr(total), ptr, nbytes, self._stream()), "b2_filter_lt_64_dev")
        return out, batch_end, total

    def aggr_dev(self, col, valid=None, out=None, dtype=np.uint32):
        """sum / count / min / max of the valid rows in one pass; returns a 3 x int64 device tensor
        laid out as b2_aggr_u32 (decode with :func:`decode_aggr`). dtype: uint32 or int32; uint64 /
        int64 (col = int64 tensor): 4 x int64 laid out as b2_aggr_u64."""
        import torch
        if np.dtype(dtype) in _DTYPES64:
            if out is None:
                out = torch.empty(4, dtype=torch.int64, device=col.device)
            self._ck(self._lib.b2_aggr_64_dev(self._h, _dptr(col), _DTYPES64[np.dtype(dtype)], _dptr(valid),
                                              col.numel(), _dptr(out), self._stream()), "b2_aggr_64_dev")
            return out
        if out is None:
            out = torch.empty(3, dtype=torch.int64, device=col.device)
        self._ck(self._lib.b2_aggr_32_dev(self._h, _dptr(col), _DTYPES32[np.dtype(dtype)], _dptr(valid),
                                          col.numel(), _dptr(out), self._stream()), "b2_aggr_32_dev")
        return out

    def take_nullable_dev(self, values, values_valid, values_len: int, indices, indices_valid, idx_len: int,
                          nbatches: int, out=None, out_valid=None):
        """Returns (out, out_valid bitmap as uint8 device tensor)."""
        import torch
        n = nbatches * idx_len
        if out is None:
            out = torch.empty(max(n, 1), dtype=torch.int32, device=values.device)
        if out_valid is None:
            out_valid = torch.zeros(((n + 31) // 32) * 4 + 4, dtype=torch.uint8, device=values.device)
        self._ck(self._lib.b2_take_u32_nullable_dev(self._h, _dptr(values), _dptr(values_valid), values_len,
                                                    _dptr(indices), _dptr(indices_valid), idx_len, nbatches,
                                                    _dptr(out), _dptr(out_valid), self._stream()),
                 "b2_take_u32_nullable_dev")
        return out[:n], out_valid

    def filter_ragged_dev(self, col, batch_off: np.ndarray, threshold: int):
        import torch
        dev = col.device
        off = np.ascontiguousarray(np.asarray(batch_off, dtype=np.int64))
        nb = off.size - 1
        d_off = torch.from_numpy(off).to(dev)
        out = torch.empty(max(int(off[-1]), 1), dtype=torch.int32, device=dev)
        batch_end = torch.empty(max(nb, 1), dtype=torch.int64, device=dev)
        total = torch.empty(1, dtype=torch.int64, device=dev)
        p_off = off.ctypes.data_as(C.POINTER(C.c_int64))
        need = int(self._lib.b2_filter_ragged_ws_bytes(p_off, nb))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._ck(self._lib.b2_filter_lt_u32_ragged_dev(self._h, _dptr(col), p_off, _dptr(d_off), nb,
                                                       threshold, _dptr(out), _dptr(batch_end),
                                                       _dptr(total), 0, _dptr(ws), ws.numel(),
                                                       self._stream()),
                 "b2_filter_lt_u32_ragged_dev")
        return out, batch_end, total

    def take_dev(self, values, values_len: int, indices, idx_len: int, nbatches: int, out=None):
        import torch
        if out is None:
            out = torch.empty(nbatches * idx_len, dtype=torch.int32, device=values.device)
        self._ck(self._lib.b2_take_u32_dev(self._h, _dptr(values), values_len, _dptr(indices), idx_len,
                                           nbatches, _dptr(out), self._stream()), "b2_take_u32_dev")
        return out

    def take64_dev(self, values, values_len: int, indices, idx_len: int, nbatches: int, out=None):
        """Batch-local gather over 64-bit values (int64 tensor of raw words), 32-bit indices."""
        import torch
        if out is None:
            out = torch.empty(nbatches * idx_len, dtype=torch.int64, device=values.device)
        self._ck(self._lib.b2_take_64_dev(self._h, _dptr(values), values_len, _dptr(indices), idx_len,
                                          nbatches, _dptr(out), self._stream()), "b2_take_64_dev")
        return out

    def take_ragged_dev(self, values, values_off: np.ndarray, indices, idx_off: np.ndarray):
        import torch
        dev = values.device
        voff = torch.from_numpy(np.ascontiguousarray(values_off, dtype=np.int64)).to(dev)
        ioff = torch.from_numpy(np.ascontiguousarray(idx_off, dtype=np.int64)).to(dev)
        nb = len(idx_off) - 1
        n = int(idx_off[-1])
        out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self._ck(self._lib.b2_take_u32_ragged_dev(self._h, _dptr(values), _dptr(voff), _dptr(indices),
                                                  _dptr(ioff), nb, 0, n, _dptr(out), self._stream()),
                 "b2_take_u32_ragged_dev")
        return out[:n]

    def partition_dev(self, cols: Sequence, nparts: int, skip_bits: int = 0):
        """cols[0] is the key column. Returns (partitioned cols, part_off[nparts+1])."""
        import torch
        dev = cols[0].device
        n = cols[0].numel()
        outs = [torch.empty(max(n, 1), dtype=torch.int32, device=dev) for _ in cols]
        off = torch.empty(nparts + 1, dtype=torch.int64, device=dev)
        need = int(self._lib.b2_partition_ws_bytes(n, nparts))
        ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (_dptr(ws) + 255) // 256 * 256
        pin = (C.c_void_p * len(cols))(*[_dptr(c) for c in cols])
        pout = (C.c_void_p * len(cols))(*[_dptr(c) for c in outs])
        self._ck(self._lib.b2_partition_u32_dev(self._h, pin, pout, len(cols), n, nparts, skip_bits,
                                                _dptr(off), ws_ptr, need, self._stream()),
                 "b2_partition_u32_dev")
        return [o[:n] for o in outs], off

    def join_ws_bytes(self, nl: int, nr: int) -> int:
        return int(self._lib.b2_join_ws_bytes(nl, nr))

    def join_min_ws_bytes(self, nl: int, nr: int) -> int:
        return int(self._lib.b2_join_min_ws_bytes(nl, nr))

    def join_ws_bytes_adjacent_outputs(self, nl: int, nr: int) -> int:
        """One-go workspace when the three output columns are one allocation (see b200olap.h)."""
        return int(self._lib.b2_join_ws_bytes_adjacent_outputs(nl, nr))

    @staticmethod
    def join_rows(out_rows) -> int:
        """Reads the device row counter of a join; ~0 (a hash-space slice overflowed its buffer:
        sliced workspace + heavily skewed keys) is an error, not a row count."""
        n = int(out_rows.cpu().numpy().view("uint64")[0])
        if n == 0xFFFFFFFFFFFFFFFF:
            raise B2Error(5, "join", "a hash-space slice overflowed; give the join a larger workspace")
        return n

    def join_dev(self, fk, y, pk, x, out_capacity: int | None = None, ws=None, outs=None,
                 out_rows=None, skip_bits: int = 0):
        """Returns (out_fk, out_y, out_x, out_rows[1] uint64-as-int64 device tensor)."""
        import torch
        dev = fk.device
        nl, nr = fk.numel(), pk.numel()
        cap = nl if out_capacity is None else int(out_capacity)
        if outs is None:
            outs = [torch.empty(max(cap, 1), dtype=torch.int32, device=dev) for _ in range(3)]
        if out_rows is None:
            out_rows = torch.empty(1, dtype=torch.int64, device=dev)
        if ws is None:
            ws = torch.empty(self.join_ws_bytes(nl, nr) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (_dptr(ws) + 255) // 256 * 256
        ws_bytes = ws.numel() - (ws_ptr - _dptr(ws))
        self._ck(self._lib.b2_join_u32_dev(self._h, _dptr(fk), _dptr(y), nl, _dptr(pk), _dptr(x), nr,
                                           _dptr(outs[0]), _dptr(outs[1]), _dptr(outs[2]), cap,
                                           _dptr(out_rows), skip_bits, ws_ptr, ws_bytes, self._stream()),
                 "b2_join_u32_dev")
        return outs[0], outs[1], outs[2], out_rows

    def join_aggr_dev(self, fk, y, pk, x, y_threshold: int | None = None, ws=None, out=None, skip_bits: int = 0):
        """Fused [filter L.y < y_threshold ->] join -> aggregate (b2_join_aggr_u32_dev): returns a
        3 x int64 device tensor (rows, sum_y, sum_x as uint64 bit patterns); nothing is materialised."""
        import torch
        nl, nr = fk.numel(), pk.numel()
        if out is None:
            out = torch.empty(3, dtype=torch.int64, device=fk.device)
        if ws is None:
            ws = torch.empty(self.join_ws_bytes(nl, nr) + 256, dtype=torch.uint8, device=fk.device)
        ws_ptr = (_dptr(ws) + 255) // 256 * 256
        ws_bytes = ws.numel() - (ws_ptr - _dptr(ws))
        self._ck(self._lib.b2_join_aggr_u32_dev(self._h, _dptr(fk), _dptr(y), nl, _dptr(pk), _dptr(x), nr,
                                                0 if y_threshold is None else 1,
                                                0 if y_threshold is None else int(y_threshold), _dptr(out),
                                                skip_bits, ws_ptr, ws_bytes, self._stream()),
                 "b2_join_aggr_u32_dev")
        return out

    def join_pairs_dev(self, l_pairs, r_pairs, out_capacity: int, skip_bits: int, ws=None, outs=None,
                       out_rows=None):
        import torch
        dev = l_pairs.device
        nl, nr = l_pairs.numel(), r_pairs.numel()
        cap = int(out_capacity)
        if outs is None:
            outs = [torch.empty(max(cap, 1), dtype=torch.int32, device=dev) for _ in range(3)]
        if out_rows is None:
            out_rows = torch.empty(1, dtype=torch.int64, device=dev)
        if ws is None:
            ws = torch.empty(self.join_ws_bytes(nl, nr) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (_dptr(ws) + 255) // 256 * 256
        ws_bytes = ws.numel() - (ws_ptr - _dptr(ws))
        self._ck(self._lib.b2_join_pairs_dev(self._h, _dptr(l_pairs), nl, _dptr(r_pairs), nr,
                                             _dptr(outs[0]), _dptr(outs[1]), _dptr(outs[2]), cap,
                                             _dptr(out_rows), skip_bits, ws_ptr, ws_bytes, self._stream()),
                 "b2_join_pairs_dev")
        return outs[0], outs[1], outs[2], out_rows

    def shuffle_partition_dev(self, key, val, nranks: int, pairs_out=None, dest_off=None, ws=None):
        """Route (key, val) rows by destination rank. Returns (pairs int64[n], dest_off int64[nranks+1])."""
        import torch
        dev = key.device
        n = key.numel()
        if pairs_out is None:
            pairs_out = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        if dest_off is None:
            dest_off = torch.empty(nranks + 1, dtype=torch.int64, device=dev)
        need = int(self._lib.b2_shuffle_ws_bytes(n, nranks))
        if ws is None:
            ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (_dptr(ws) + 255) // 256 * 256
        self._ck(self._lib.b2_shuffle_partition_u32_dev(self._h, _dptr(key), _dptr(val), n, nranks,
                                                        _dptr(pairs_out), _dptr(dest_off), ws_ptr,
                                                        ws.numel() - (ws_ptr - _dptr(ws)), self._stream()),
                 "b2_shuffle_partition_u32_dev")
        return pairs_out[:n], dest_off

    # ---- fused multi-GPU shuffle over peer memory --------------------------------------------------
    def shuffle_p2p_ws_bytes(self, n: int, bits: int) -> int:
        return int(self._lib.b2_shuffle_p2p_ws_bytes(n, bits))

    @staticmethod
    def _aligned(ws):
        ptr = (_dptr(ws) + 255) // 256 * 256
        return ptr, ws.numel() - (ptr - _dptr(ws))

    def shuffle_p2p_count_dev(self, key, bits: int, ws, bucket_off=None, val=None, val_lt: int | None = None):
        """Rows per bucket (top `bits` bits of wang_hash): returns int64[2^bits + 1] boundaries. val_lt: only
        rows with val < val_lt are counted (b2_shuffle_p2p_count_lt_dev: a predicate pushed in front of the link)."""
        import torch
        if bucket_off is None:
            bucket_off = torch.empty((1 << bits) + 1, dtype=torch.int64, device=key.device)
        ptr, nbytes = self._aligned(ws)
        if val_lt is None:
            self._ck(self._lib.b2_shuffle_p2p_count_dev(self._h, _dptr(key), key.numel(), bits, _dptr(bucket_off),
                                                        ptr, nbytes, self._stream()), "b2_shuffle_p2p_count_dev")
        else:
            self._ck(self._lib.b2_shuffle_p2p_count_lt_dev(self._h, _dptr(key), _dptr(val), key.numel(), bits, 1,
                                                           int(val_lt), _dptr(bucket_off), ptr, nbytes, self._stream()),
                     "b2_shuffle_p2p_count_lt_dev")
        return bucket_off

    def shuffle_p2p_scatter_dev(self, key, val, bits: int, bucket_addr, ws, abort=None, val_lt: int | None = None):
        """Writes the (key, val) pairs of bucket b contiguously from byte address bucket_addr[b]
        (int64 device tensor of 2^bits addresses — local or peer memory). abort: int64[1] device
        tensor; non-zero on the device = store nothing (the plan kernel's overflow flag). val_lt: only rows
        with val < val_lt are sent (the count must have used the same predicate)."""
        ptr, nbytes = self._aligned(ws)
        self._ck(self._lib.b2_shuffle_p2p_scatter_lt_dev(self._h, _dptr(key), _dptr(val), key.numel(), bits,
                                                         0 if val_lt is None else 1, 0 if val_lt is None else int(val_lt),
                                                         _dptr(bucket_addr), None if abort is None else _dptr(abort),
                                                         ptr, nbytes, self._stream()),
                 "b2_shuffle_p2p_scatter_lt_dev")

    def shuffle_p2p_plan_dev(self, off_ptrs, recv_base, rank: int, nranks: int, bits: int, capacity_rows: int,
                             bucket_addr, seg_off, info, prev_abort=None):
        """b2_shuffle_p2p_plan_dev: every rank's bucket boundaries -> this rank's destination addresses,
        the coarse-bucket boundaries it receives and {rows received, max over ranks, overflow} — one
        launch, nothing read back. off_ptrs / recv_base: int64 device tensors of nranks entries."""
        self._ck(self._lib.b2_shuffle_p2p_plan_dev(self._h, _dptr(off_ptrs), _dptr(recv_base), rank, nranks, bits,
                                                   int(capacity_rows), _dptr(bucket_addr), _dptr(seg_off),
                                                   _dptr(info), None if prev_abort is None else _dptr(prev_abort),
                                                   self._stream()), "b2_shuffle_p2p_plan_dev")

    def join_seg_cap_ws_bytes(self, nl_cap: int, nr_cap: int, nr_expected: int, skip_bits: int, seg_bits: int) -> int:
        return int(self._lib.b2_join_seg_cap_ws_bytes(nl_cap, nr_cap, nr_expected, skip_bits, seg_bits))

    def join_pairs_seg_cap_dev(self, l_pairs, l_seg_off, r_pairs, r_seg_off, nr_expected: int, seg_bits: int,
                               out_capacity: int, skip_bits: int, ws, outs, out_rows, abort=None, phases: int = 7):
        """b2_join_pairs_seg_cap_phased_dev: l_pairs / r_pairs are whole receive BUFFERS (their sizes are
        capacities); the rows really there are the last entries of the segment tables, on the device.
        phases: 1 build | 2 probe (repeatable, one share of the probe side per call) | 4 finish."""
        ptr, nbytes = self._aligned(ws)
        self._ck(self._lib.b2_join_pairs_seg_cap_phased_dev(
            self._h, _dptr(l_pairs), _dptr(l_seg_off), l_pairs.numel(), _dptr(r_pairs), _dptr(r_seg_off),
            r_pairs.numel(), int(nr_expected), seg_bits, _dptr(outs[0]), _dptr(outs[1]), _dptr(outs[2]),
            int(out_capacity), _dptr(out_rows), skip_bits, None if abort is None else _dptr(abort), int(phases),
            ptr, nbytes, self._stream()), "b2_join_pairs_seg_cap_phased_dev")
        return outs[0], outs[1], outs[2], out_rows

    def join_aggr_pairs_seg_cap_dev(self, l_pairs, l_seg_off, r_pairs, r_seg_off, nr_expected: int, seg_bits: int,
                                    skip_bits: int, ws, out, y_threshold: int | None = None, abort=None,
                                    phases: int = 7):
        """b2_join_aggr_pairs_seg_cap_phased_dev: the fused join -> aggregate pipeline over the receive buffers
        of the shuffle (arguments as join_pairs_seg_cap_dev); out = 3 x int64 device tensor (rows, sum of the
        left payloads, sum of the right payloads as uint64 bit patterns), written by phases & 4."""
        ptr, nbytes = self._aligned(ws)
        self._ck(self._lib.b2_join_aggr_pairs_seg_cap_phased_dev(
            self._h, _dptr(l_pairs), _dptr(l_seg_off), l_pairs.numel(), _dptr(r_pairs), _dptr(r_seg_off),
            r_pairs.numel(), int(nr_expected), seg_bits, 0 if y_threshold is None else 1,
            0 if y_threshold is None else int(y_threshold), _dptr(out), skip_bits,
            None if abort is None else _dptr(abort), int(phases), ptr, nbytes, self._stream()),
            "b2_join_aggr_pairs_seg_cap_phased_dev")
        return out

    def join_seg_ws_bytes(self, nl: int, nr: int, skip_bits: int, seg_bits: int) -> int:
        return int(self._lib.b2_join_seg_ws_bytes(nl, nr, skip_bits, seg_bits))

    def join_pairs_seg_dev(self, l_pairs, l_seg_off, r_pairs, r_seg_off, seg_bits: int, out_capacity: int,
                           skip_bits: int, ws=None, outs=None, out_rows=None):
        """Join of sides already grouped into 2^seg_bits coarse buckets (see b2_join_pairs_seg_dev)."""
        import torch
        dev = l_pairs.device
        nl, nr = l_pairs.numel(), r_pairs.numel()
        cap = int(out_capacity)
        if outs is None:
            outs = [torch.empty(max(cap, 1), dtype=torch.int32, device=dev) for _ in range(3)]
        if out_rows is None:
            out_rows = torch.empty(1, dtype=torch.int64, device=dev)
        if ws is None:
            ws = torch.empty(self.join_seg_ws_bytes(nl, nr, skip_bits, seg_bits) + 256, dtype=torch.uint8,
                             device=dev)
        ptr, nbytes = self._aligned(ws)
        self._ck(self._lib.b2_join_pairs_seg_dev(self._h, _dptr(l_pairs), _dptr(l_seg_off), nl, _dptr(r_pairs),
                                                 _dptr(r_seg_off), nr, seg_bits, _dptr(outs[0]), _dptr(outs[1]),
                                                 _dptr(outs[2]), cap, _dptr(out_rows), skip_bits, ptr, nbytes,
                                                 self._stream()), "b2_join_pairs_seg_dev")
        return outs[0], outs[1], outs[2], out_rows


class AggrResult(C.Structure):
    """b2_aggr_u32 (include/b200olap.h)."""
    _fields_ = [("sum", C.c_uint64), ("count", C.c_uint64), ("min", C.c_uint32), ("max", C.c_uint32)]

    def as_dict(self, dtype=np.uint32) -> dict:
        """Arrow's scalars: every aggregate of zero valid rows is null (None); count never is.
        int32 columns: the fields hold int64 / int32 bit patterns."""
        empty = self.count == 0
        sgn = np.dtype(dtype) == np.dtype(np.int32)
        s = int(self.sum) - (1 << 64) if sgn and self.sum >> 63 else int(self.sum)
        lo = int(self.min) - (1 << 32) if sgn and self.min >> 31 else int(self.min)
        hi = int(self.max) - (1 << 32) if sgn and self.max >> 31 else int(self.max)
        return {"sum": None if empty else s, "count": int(self.count),
                "min": None if empty else lo, "max": None if empty else hi}


class AggrResult64(C.Structure):
    """b2_aggr_u64 (include/b200olap.h)."""
    _fields_ = [("sum", C.c_uint64), ("count", C.c_uint64), ("min", C.c_uint64), ("max", C.c_uint64)]

    def as_dict(self, dtype=np.uint64) -> dict:
        empty = self.count == 0
        sgn = np.dtype(dtype) == np.dtype(np.int64)
        dec = lambda v: int(v) - (1 << 64) if sgn and v >> 63 else int(v)
        return {"sum": None if empty else dec(self.sum), "count": int(self.count),
                "min": None if empty else dec(self.min), "max": None if empty else dec(self.max)}


def decode_aggr(t, dtype=np.uint32) -> dict:
    """Decode the device tensor returned by :meth:`Context.aggr_dev`."""
    raw = t.cpu().numpy().tobytes()
    if np.dtype(dtype) in _DTYPES64:
        return AggrResult64.from_buffer_copy(raw[:C.sizeof(AggrResult64)]).as_dict(dtype)
    return AggrResult.from_buffer_copy(raw[:C.sizeof(AggrResult)]).as_dict(dtype)


def wang_hash(key: int) -> int:
    return int(_lib.lib().b2_wang_hash_u32(int(key) & 0xFFFFFFFF))


def join_dest_rank(key: int, nranks: int) -> int:
    return int(_lib.lib().b2_join_dest_rank(int(key) & 0xFFFFFFFF, int(nranks)))


# ---------------------------------------------------------------------------------------------
# operators (host batches in, host results out) — the reference's *Dpu classes
# ---------------------------------------------------------------------------------------------
class DeviceSet:
    """The GPUs of one node behind one handle (b2_set): what ``dpu::DpuSet::allocate(nr_dpus)``
    (dpuext.hpp:704-739) is to the reference's operators. Pass it to FilterGpu / SumGpu / TakeGpu /
    JoinGpu in place of a Context: filter, sum and take shard by batch range, the join runs the fused
    peer-memory shuffle over NVLink — all inside libb200olap.so, one host process, no torch."""

    _SET_OPS = {"sum_u32_host", "filter_lt_u32_host", "filter_fetch_host", "take_u32_host", "join_u32_host",
                "join_fetch_host", "join_aggr_u32_host"}

    def __init__(self, devices: Sequence[int] | int | None = None):
        self._lib = _lib.lib()
        if devices is None:
            n = C.c_int(0)
            check(self._lib.b2_device_count(C.byref(n)), "b2_device_count")
            devices = list(range(n.value))
        elif isinstance(devices, int):
            devices = list(range(devices))
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(self._lib.b2_set_create(arr, len(devices), C.byref(h)), "b2_set_create")
        self._h = h
        self.devices = list(devices)
        self.members = [Context(_borrowed=self._lib.b2_set_ctx(h, i)) for i in range(len(devices))]

    def __len__(self) -> int:
        return len(self.devices)

    @property
    def peer_access(self) -> bool:
        return bool(self._lib.b2_set_peer_access(self._h))

    @property
    def launches(self) -> int:
        return int(self._lib.b2_set_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            for m in self.members:
                m.close()
            self._lib.b2_set_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, status: int, where: str) -> None:
        if status != _lib.B2_OK:
            raise B2Error(status, where, self._lib.b2_set_last_error(self._h).decode(errors="replace"))

    def join_trace(self, on: bool = True) -> None:
        for m in self.members:
            m.join_trace(on)

    def join_last_phases(self) -> dict:
        """The slowest member per phase: the members run their local joins side by side."""
        per = [m.join_last_phases() for m in self.members]
        return {k: max(p[k] for p in per) for k in per[0]}

    def _call(self, name: str, *args) -> None:
        """b2_set_<name>(set, ...) where the set shards the operator; otherwise (nullable / typed /
        64-bit variants) the entry point of member 0."""
        if name in self._SET_OPS:
            self._ck(getattr(self._lib, "b2_set_" + name)(self._h, *args), "b2_set_" + name)
        else:
            self.members[0]._call(name, *args)


class ArrowArrayStruct(C.Structure):
    """struct ArrowArray of the Arrow C Data Interface (arrow/c/abi.h)."""


ArrowArrayStruct._fields_ = [
    ("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
    ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)), ("children", C.c_void_p),
    ("dictionary", C.c_void_p), ("release", C.CFUNCTYPE(None, C.POINTER(ArrowArrayStruct))),
    ("private_data", C.c_void_p)]


class ArrowDeviceArrayStruct(C.Structure):
    """struct ArrowDeviceArray of the Arrow C Device Data Interface."""
    _fields_ = [("array", ArrowArrayStruct), ("device_id", C.c_int64), ("device_type", C.c_int32),
                ("sync_event", C.c_void_p), ("reserved", C.c_int64 * 3)]


ARROW_DEVICE_CUDA = 2


class DeviceColumn:
    """A packed uint32 column resident in HBM (b2_col): operators over DeviceColumns leave their
    results on the device, so FilterGpu -> TakeGpu -> SumGpu style chains cross PCIe once on the way
    in and not in between (the reference copies in and out around every operator and lists the
    zero-copy result as future work, arrow_utils.h:28-29). export_arrow() / import_arrow() speak the
    Arrow C Device Data Interface: zero-copy hand-over to any Arrow consumer on the same GPU."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, C.c_void_p(handle)

    @classmethod
    def from_host(cls, ctx: Context, batches: Sequence[Any]) -> "DeviceColumn":
        tab = _PtrTable([_column(b, 0) if not isinstance(b, np.ndarray) else _as_u32(b) for b in batches])
        h = C.c_void_p()
        ctx._ck(ctx._lib.b2_col_upload_host(ctx._h, tab.ptrs, tab.lens, tab.n, C.byref(h)), "b2_col_upload_host")
        return cls(ctx, h.value)

    @classmethod
    def import_arrow(cls, ctx: Context, dev_array: ArrowDeviceArrayStruct, batch_lens: Sequence[int] | None = None):
        """Takes over an ArrowDeviceArray (its release callback moves to the column)."""
        n = len(batch_lens) if batch_lens else 0
        lens = (C.c_int64 * max(n, 1))(*(batch_lens or [0]))
        h = C.c_void_p()
        ctx._ck(ctx._lib.b2_col_import(ctx._h, C.byref(dev_array), lens if n else None, n, C.byref(h)), "b2_col_import")
        return cls(ctx, h.value)

    def export_arrow(self) -> ArrowDeviceArrayStruct:
        """Zero-copy ArrowDeviceArray view; the device memory lives until its release() has run."""
        a = ArrowDeviceArrayStruct()
        self.ctx._ck(self.ctx._lib.b2_col_export(self._h, C.byref(a)), "b2_col_export")
        return a

    @property
    def rows(self) -> int:
        return int(self.ctx._lib.b2_col_rows(self._h))

    @property
    def nbatches(self) -> int:
        return int(self.ctx._lib.b2_col_nbatches(self._h))

    @property
    def device_ptr(self) -> int:
        return int(self.ctx._lib.b2_col_device_ptr(self._h) or 0)

    def batch_offsets(self) -> np.ndarray:
        out = np.zeros(self.nbatches + 1, dtype=np.int64)
        self.ctx._ck(self.ctx._lib.b2_col_batch_offsets(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), out.size),
                     "b2_col_batch_offsets")
        return out

    def to_host(self) -> list[np.ndarray]:
        off = self.batch_offsets()
        outs = [np.empty(int(off[b + 1] - off[b]), dtype=np.uint32) for b in range(self.nbatches)]
        ptrs = (C.c_void_p * max(self.nbatches, 1))(*[o.ctypes.data for o in outs])
        self.ctx._ck(self.ctx._lib.b2_col_download_host(self.ctx._h, self._h, ptrs, self.nbatches), "b2_col_download_host")
        return outs

    # ---- operators: results stay on the device ----
    def sum(self) -> int:
        out = C.c_uint64(0)
        self.ctx._ck(self.ctx._lib.b2_sum_u32_col(self.ctx._h, self._h, C.byref(out)), "b2_sum_u32_col")
        return int(out.value)

    def filter_lt(self, threshold: int = 1 << 30) -> "DeviceColumn":
        h = C.c_void_p()
        self.ctx._ck(self.ctx._lib.b2_filter_lt_u32_col(self.ctx._h, self._h, int(threshold), C.byref(h)),
                     "b2_filter_lt_u32_col")
        return DeviceColumn(self.ctx, h.value)

    def take(self, indices: "DeviceColumn") -> "DeviceColumn":
        h = C.c_void_p()
        self.ctx._ck(self.ctx._lib.b2_take_u32_col(self.ctx._h, self._h, indices._h, C.byref(h)), "b2_take_u32_col")
        return DeviceColumn(self.ctx, h.value)

    @staticmethod
    def join(fk: "DeviceColumn", y: "DeviceColumn", pk: "DeviceColumn", x: "DeviceColumn"):
        ctx = fk.ctx
        hs = [C.c_void_p() for _ in range(3)]
        ctx._ck(ctx._lib.b2_join_u32_col(ctx._h, fk._h, y._h, pk._h, x._h, *[C.byref(h) for h in hs]), "b2_join_u32_col")
        return tuple(DeviceColumn(ctx, h.value) for h in hs)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.ctx._lib.b2_col_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


FILTER_THRESHOLD = 1 << 30  # predicate v < 2^30: filter.c:25, filter_native.cc:59


class FilterGpu:
    """FilterDpu (host/filter/filter_dpu.cc:23-174): order-preserving ``v < threshold``."""

    def __init__(self, ctx: Context, batches: Sequence[Any], threshold: int = FILTER_THRESHOLD):
        self.ctx = ctx
        self._ncols = [_nullable_column(b, 0, typed=True, wide=True, f64=True) for b in batches]
        self._cols = [c.values for c in self._ncols]
        self._valid = _ValidTable(self._ncols)
        kinds = {c.dtype for c in self._ncols}
        if len(kinds) > 1:
            raise TypeError(f"batches of different types: {sorted(str(k) for k in kinds)}")
        # uint32 as the reference; int32 / float32 and uint64 / int64 / float64 too
        self.dtype = kinds.pop() if kinds else np.dtype(np.uint32)
        self.threshold = threshold if self.dtype.kind == "f" else int(threshold)
        self._timers = None

    def Prepare(self) -> None:  # the reference loads the DPU binary here (filter_dpu.cc:23-32)
        self._timers = Timers()

    def _run(self):
        tab = _PtrTable(self._cols)
        counts = (C.c_int64 * max(tab.n, 1))()
        total = C.c_uint64(0)
        t1 = Timings()
        self.ctx._call("filter_lt_u32_host", tab.ptrs, tab.lens, tab.n, self.threshold, counts, C.byref(total),
                       C.byref(t1))
        return tab, counts, total.value, t1

    def Run(self) -> int:
        """Number of selected rows (FilterDpu::Run, filter_dpu.cc:171-173)."""
        return self.GetResult(_count_only=True)

    def _get_result_nullable(self, _count_only: bool):
        """Nullable input: a row is kept iff valid and below the threshold (Arrow's filter drops
        null predicates); the result has no nulls."""
        tab = _PtrTable(self._cols)
        counts = (C.c_int64 * max(tab.n, 1))()
        total = C.c_uint64(0)
        flat = np.empty(sum(a.size for a in self._cols), dtype=self.dtype)
        t = Timings()
        if self.dtype.itemsize == 8:  # counted two-pass compaction (csrc/filter64.cu)
            self.ctx._call("filter_lt_64_host_into", tab.ptrs, self._valid.ptrs, self._valid.offs, tab.lens, tab.n,
                           _DTYPES64_FILTER[self.dtype], _threshold_bits(self.threshold, self.dtype), flat.ctypes.data,
                           flat.size, counts, C.byref(total), C.byref(t))
        else:
            self.ctx._call("filter_lt_32_host_into", tab.ptrs, self._valid.ptrs, self._valid.offs, tab.lens, tab.n,
                           _DTYPES32[self.dtype], _threshold_bits(self.threshold, self.dtype), flat.ctypes.data,
                           flat.size, counts, C.byref(total), C.byref(t))
        self._timers = Timers.from_timings(t)
        self._last = (t,)
        if _count_only:
            return int(total.value)
        bounds = np.concatenate([[0], np.cumsum(np.frombuffer(counts, dtype=np.int64, count=tab.n))])
        return [flat[bounds[b]:bounds[b + 1]] for b in range(tab.n)]

    def GetResult(self, _count_only: bool = False):
        """One uint32 array per input batch, in batch order (ChunkedArray chunks, :162-166)."""
        if self._valid.any or self.dtype != np.dtype(np.uint32):
            return self._get_result_nullable(_count_only)
        tab, counts, total, t1 = self._run()
        # the reference's Run() is GetResult()->length(): the result is always pulled back
        flat = np.empty(total, dtype=np.uint32)
        ptrs = (C.c_void_p * max(tab.n, 1))()
        off = 0
        for b in range(tab.n):
            ptrs[b] = flat.ctypes.data + 4 * off
            off += counts[b]
        t2 = Timings()
        self.ctx._call("filter_fetch_host", ptrs, tab.n, C.byref(t2))
        self._timers = Timers.from_timings(t1, t2)
        self._last = (t1, t2)
        if _count_only:
            return int(total)
        bounds = np.concatenate([[0], np.cumsum(np.frombuffer(counts, dtype=np.int64, count=tab.n))])
        return [flat[bounds[b]:bounds[b + 1]] for b in range(tab.n)]

    def Timers(self):
        return self._timers


class SumGpu:
    """SumDpu (host/aggr/aggr_dpu.cc:31-89): sum of a uint32 column as uint64. Also int32 (sum as
    int64) and uint64 / int64 columns (sum in the column's type, wrapping as Arrow's does)."""

    def __init__(self, ctx: Context, batches: Sequence[Any]):
        self.ctx = ctx
        self._ncols = [_nullable_column(b, 0, typed=True, wide=True) for b in batches]
        self._cols = [c.values for c in self._ncols]
        self._valid = _ValidTable(self._ncols)
        kinds = {c.dtype for c in self._ncols} or {np.dtype(np.uint32)}
        if len(kinds) > 1 or next(iter(kinds)) == np.dtype(np.float32):
            raise TypeError(f"aggregates take uint32 / int32 / uint64 / int64 columns, got {sorted(str(k) for k in kinds)}")
        self.dtype = kinds.pop()
        self._timers = None

    def Prepare(self) -> None:
        self._timers = Timers()

    def Aggregates(self) -> dict:
        """sum / count / min / max over the valid rows (Arrow semantics: None when no row is valid)."""
        tab = _PtrTable(self._cols)
        t = Timings()
        if self.dtype in _DTYPES64:
            out = AggrResult64()
            self.ctx._call("aggr_64_host", tab.ptrs, self._valid.ptrs, self._valid.offs, tab.lens, tab.n,
                           _DTYPES64[self.dtype], C.byref(out), C.byref(t))
        else:
            out = AggrResult()
            self.ctx._call("aggr_32_host", tab.ptrs, self._valid.ptrs, self._valid.offs, tab.lens, tab.n,
                           _DTYPES32[self.dtype], C.byref(out), C.byref(t))
        self._timers = Timers.from_timings(t)
        self._last = (t,)
        return out.as_dict(self.dtype)

    def Run(self):
        if self._valid.any or self.dtype != np.dtype(np.uint32):
            # nullable / int32 column: Arrow's sum skips nulls, and is null if all rows are
            return self.Aggregates()["sum"]
        tab = _PtrTable(self._cols)
        out = C.c_uint64(0)
        t = Timings()
        self.ctx._call("sum_u32_host", tab.ptrs, tab.lens, tab.n, C.byref(out), C.byref(t))
        self._timers = Timers.from_timings(t)
        self._last = (t,)
        return int(out.value)

    def Timers(self):
        return self._timers


class TakeGpu:
    """TakeDpu (host/take/take_dpu.cc:34-104): batch-local gather, no bounds check."""

    def __init__(self, ctx: Context, batches: Sequence[Any], indices_batches: Sequence[Any]):
        if len(batches) != len(indices_batches):
            raise ValueError("values and indices must have the same number of batches")
        self.ctx = ctx
        self._nvals = [_nullable_column(b, 0, wide=True) for b in batches]  # uint32, or 64-bit values
        self._nidx = [_nullable_column(b, 0) for b in indices_batches]
        self._vals = [c.values for c in self._nvals]
        self._idx = [c.values for c in self._nidx]
        self._vvalid, self._ivalid = _ValidTable(self._nvals), _ValidTable(self._nidx)
        kinds = {c.dtype for c in self._nvals}
        if len(kinds) > 1:
            raise TypeError(f"value batches of different types: {sorted(str(k) for k in kinds)}")
        self.dtype = kinds.pop() if kinds else np.dtype(np.uint32)
        if self.dtype in _DTYPES64 and (self._vvalid.any or self._ivalid.any):
            raise TypeError("take over 64-bit values handles non-null columns only")
        self._timers = None

    def Prepare(self) -> None:
        self._timers = Timers()

    def _run_nullable(self):
        """Nullable values and / or indices: one pyarrow uint32 array per batch; slot j is null when
        index j is null or the value it selects is null (cp::Take)."""
        v, i = _PtrTable(self._vals), _PtrTable(self._idx)
        outs = [np.empty(a.size, dtype=np.uint32) for a in self._idx]
        bits = [np.zeros((a.size + 7) // 8, dtype=np.uint8) for a in self._idx]
        optrs = (C.c_void_p * max(i.n, 1))(*[o.ctypes.data for o in outs])
        bptrs = (C.c_void_p * max(i.n, 1))(*[b.ctypes.data for b in bits])
        t = Timings()
        self.ctx._call("take_u32_nullable_host", v.ptrs, self._vvalid.ptrs, self._vvalid.offs, v.lens, i.ptrs,
                       self._ivalid.ptrs, self._ivalid.offs, i.lens, i.n, optrs, bptrs, C.byref(t))
        self._timers = Timers.from_timings(t)
        self._last = (t,)
        return [_to_arrow(o, b) for o, b in zip(outs, bits)]

    def Run(self):
        """One uint32 array per batch (the reference returns a Table of one chunk per batch); 64-bit
        value batches give arrays of their own type."""
        if self._vvalid.any or self._ivalid.any:
            return self._run_nullable()
        v, i = _PtrTable(self._vals), _PtrTable(self._idx)
        total = sum(a.size for a in self._idx)
        wide = self.dtype in _DTYPES64
        flat = np.empty(total, dtype=self.dtype if wide else np.uint32)
        ptrs = (C.c_void_p * max(i.n, 1))()
        bounds = [0]
        for b, a in enumerate(self._idx):
            ptrs[b] = flat.ctypes.data + flat.itemsize * bounds[-1]
            bounds.append(bounds[-1] + a.size)
        t = Timings()
        if wide:
            self.ctx._call("take_64_host", v.ptrs, v.lens, i.ptrs, i.lens, i.n, ptrs, C.byref(t))
        else:
            self.ctx._call("take_u32_host", v.ptrs, v.lens, i.ptrs, i.lens, i.n, ptrs, C.byref(t))
        self._timers = Timers.from_timings(t)
        self._last = (t,)
        return [flat[bounds[b]:bounds[b + 1]] for b in range(i.n)]

    def Timers(self):
        return self._timers


class JoinGpu:
    """JoinDpu (host/join/join_dpu.cc:144-400): inner join L.fk = R.pk, output fk + every other
    column of both sides (JoinNative drops pk, join_native.cc:75).

    left batches have columns (fk_name, left payloads...), right batches (pk_name, right payloads...);
    payloads are "the other columns", whatever they are called (y / x in the benchmark, v_l / v_r in
    JoinTest.SimpleTest, join_test.cc:45-64). One payload per side is the fast path (the payload
    travels with the key); any other number goes through row numbers + the take kernel, as JoinDpu's
    selection vector + TakeKernel does (join_dpu.cc:127-138,325-341).
    """

    def __init__(self, ctx, left_batches: Sequence[Any], right_batches: Sequence[Any],
                 fk: str = "fk", pk: str = "pk"):
        self.ctx = ctx
        self.fk, self.pk = fk, pk
        ln = _column_names(left_batches[0]) if len(left_batches) else [fk, "y"]
        rn = _column_names(right_batches[0]) if len(right_batches) else [pk, "x"]
        if fk not in ln or pk not in rn:
            raise ValueError(f"join keys {fk!r} / {pk!r} not among the columns {ln} / {rn}")
        self.lpays = [c for c in ln if c != fk]
        self.rpays = [c for c in rn if c != pk]
        if set(self.lpays) & set(self.rpays) or fk in self.rpays:
            raise ValueError("left and right payload columns must have distinct names")
        self.lpay = self.lpays[0] if len(self.lpays) == 1 else None
        self.rpay = self.rpays[0] if len(self.rpays) == 1 else None
        # any column that is not uint32 (64-bit keys, HT_64BIT_KEYS hashtable.h:14-18; int / float /
        # 64-bit payloads) sends the join through the typed-table entry point
        self._typed = None
        if len(left_batches) and len(right_batches):
            ldt = [_column_dtype(left_batches[0], c) for c in [fk] + self.lpays]
            rdt = [_column_dtype(right_batches[0], c) for c in [pk] + self.rpays]
            if any(d != np.dtype(np.uint32) for d in ldt + rdt):
                if ldt[0].itemsize != rdt[0].itemsize or ldt[0].kind == "f":
                    raise TypeError(f"join keys must be integers of one width, got {ldt[0]} / {rdt[0]}")
                self._typed = (ldt, rdt)
                self._l = [[_raw_column(b, c) for b in left_batches] for c in [fk] + self.lpays]
                self._r = [[_raw_column(b, c) for b in right_batches] for c in [pk] + self.rpays]
                self._timers = None
                return
        # key columns may carry nulls (a null key never matches, as in Arrow's hash join); payloads may not
        lkeys = [_nullable_column(b, fk) for b in left_batches]
        rkeys = [_nullable_column(b, pk) for b in right_batches]
        self._lkv, self._rkv = _ValidTable(lkeys), _ValidTable(rkeys)
        self._l = [[c.values for c in lkeys]] + [[_column(b, c) for b in left_batches] for c in self.lpays]
        self._r = [[c.values for c in rkeys]] + [[_column(b, c) for b in right_batches] for c in self.rpays]
        self._keep = (lkeys, rkeys)
        self._timers = None

    def Prepare(self) -> None:
        self._timers = Timers()

    def Run(self) -> dict:
        """{fk: array, left payloads..., right payloads...}, row order unspecified."""
        nlb, nrb = len(self._l[0]), len(self._r[0])
        if self._typed is None:
            lt, rt = _PtrTable([a for col in self._l for a in col]), _PtrTable([a for col in self._r for a in col])
        rows = C.c_uint64(0)
        t1, t2 = Timings(), Timings()
        names = [self.fk] + self.lpays + self.rpays
        self.ctx.join_trace(True)
        if self._typed is not None:
            ldt, rdt = self._typed

            def raw_table(cols):
                arrays = [a for col in cols for a in col]
                ptrs = (C.c_void_p * max(len(arrays), 1))(*[a.ctypes.data for a in arrays])
                lens = (C.c_int64 * max(len(cols[0]), 1))(*[int(a.size) for a in cols[0]])
                return ptrs, lens, arrays
            lp, ll, _k1 = raw_table(self._l)
            rp, rl, _k2 = raw_table(self._r)
            lb = (C.c_int * len(ldt))(*[d.itemsize for d in ldt])
            rb = (C.c_int * len(rdt))(*[d.itemsize for d in rdt])
            self.ctx._call("join_table_host", lp, ll, nlb, lb, len(ldt), rp, rl, nrb, rb, len(rdt), C.byref(rows),
                           C.byref(t1))
            n = int(rows.value)
            out = [np.empty(n, dtype=d) for d in [ldt[0]] + ldt[1:] + rdt[1:]]
            ptrs = (C.c_void_p * len(out))(*[o.ctypes.data for o in out])
            self.ctx._call("join_table_fetch_host", ptrs, len(out), n, C.byref(t2))
        elif self._lkv.any or self._rkv.any:
            if len(self.lpays) != 1 or len(self.rpays) != 1:
                raise ValueError("nullable join keys are supported with one payload column per side")
            self.ctx._call("join_u32_nullable_host", lt.ptrs, self._lkv.ptrs, self._lkv.offs, lt.lens, nlb, rt.ptrs,
                           self._rkv.ptrs, self._rkv.offs, rt.lens, nrb, C.byref(rows), C.byref(t1))
            n = int(rows.value)
            out = [np.empty(n, dtype=np.uint32) for _ in range(3)]
            self.ctx._call("join_fetch_host", out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, n,
                           C.byref(t2))
        elif len(self.lpays) == 1 and len(self.rpays) == 1:
            self.ctx._call("join_u32_host", lt.ptrs, lt.lens, nlb, rt.ptrs, rt.lens, nrb, C.byref(rows), C.byref(t1))
            n = int(rows.value)
            out = [np.empty(n, dtype=np.uint32) for _ in range(3)]
            self.ctx._call("join_fetch_host", out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, n,
                           C.byref(t2))
        else:
            self.ctx._call("join_cols_u32_host", lt.ptrs, lt.lens, nlb, len(self.lpays), rt.ptrs, rt.lens, nrb,
                           len(self.rpays), C.byref(rows), C.byref(t1))
            n = int(rows.value)
            out = [np.empty(n, dtype=np.uint32) for _ in names]
            ptrs = (C.c_void_p * len(names))(*[o.ctypes.data for o in out])
            self.ctx._call("join_cols_fetch_host", ptrs, len(names), n, C.byref(t2))
        self._timers = Timers.from_timings(t1, t2)
        self._timers.add_join_phases(self.ctx.join_last_phases())
        self.ctx.join_trace(False)
        self._last = (t1, t2)
        return dict(zip(names, out))

    def RunAggregate(self, y_threshold: int | None = None) -> dict:
        """Fused pipeline (b2_join_aggr_u32_host): COUNT(*), SUM(left payload), SUM(right payload) of
        the join, optionally over the probe rows with left payload < y_threshold only — the result
        columns are never materialised and nothing but three numbers comes back."""

        class _Aggr(C.Structure):
            _fields_ = [("rows", C.c_uint64), ("sum_y", C.c_uint64), ("sum_x", C.c_uint64)]

        if self.lpay is None or self.rpay is None or self._typed is not None:
            raise ValueError("the fused join -> aggregate pipeline takes uint32 columns, one payload per side")
        lt, rt = _PtrTable(self._l[0] + self._l[1]), _PtrTable(self._r[0] + self._r[1])
        out, t = _Aggr(), Timings()
        self.ctx.join_trace(True)
        self.ctx._call("join_aggr_u32_host", lt.ptrs, lt.lens, len(self._l[0]), rt.ptrs, rt.lens, len(self._r[0]),
                       0 if y_threshold is None else 1, 0 if y_threshold is None else int(y_threshold),
                       C.byref(out), C.byref(t))
        self._timers = Timers.from_timings(t)
        self._timers.add_join_phases(self.ctx.join_last_phases())
        self.ctx.join_trace(False)
        self._last = (t,)
        return {"rows": int(out.rows), f"sum_{self.lpay}": int(out.sum_y), f"sum_{self.rpay}": int(out.sum_x)}

    def Timers(self):
        return self._timers


class PartitionGpu:
    """PartitionDpu (host/partition/partition_dpu.h:15-28): hash-partition record batches.

    Run() returns ``nr_partitions`` record batches (dict column -> array); partition of a row =
    top log2(nr_partitions) bits of wang_hash(key) (partition.c:20-28,45-46)."""

    def __init__(self, ctx: Context, batches: Sequence[Any], nr_partitions: int, partition_key: str):
        self.ctx = ctx
        self.names = _column_names(batches[0]) if len(batches) else [partition_key]
        self.key = partition_key
        self.nparts = int(nr_partitions)
        self._cols = {n: [_column(b, n) for b in batches] for n in self.names}
        self._nb = len(batches)
        self._timers = None

    def Prepare(self) -> None:
        self._timers = Timers()

    def Run(self) -> list[dict]:
        names = self.names
        ncols = len(names)
        arrays = [a for n in names for a in self._cols[n]]  # column-major
        tab = _PtrTable(arrays)
        lens = (C.c_int64 * max(self._nb, 1))(*[int(a.size) for a in self._cols[names[0]]])
        rows = (C.c_int64 * self.nparts)()
        t1, t2 = Timings(), Timings()
        self.ctx._call("partition_u32_host", tab.ptrs, lens, self._nb, ncols, names.index(self.key), self.nparts, rows,
                       C.byref(t1))
        outs = [{n: np.empty(rows[p], dtype=np.uint32) for n in names} for p in range(self.nparts)]
        ptrs = (C.c_void_p * (self.nparts * ncols))()
        for p in range(self.nparts):
            for c, n in enumerate(names):
                ptrs[p * ncols + c] = outs[p][n].ctypes.data
        self.ctx._call("partition_fetch_host", ptrs, self.nparts, ncols, C.byref(t2))
        self._timers = Timers.from_timings(t1, t2)
        self._last = (t1, t2)
        return outs

    def Timers(self):
        return self._timers
