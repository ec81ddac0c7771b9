"""Host-side type handling of the operator classes (no GPU: the constructors only inspect their
batches): which column types each operator admits, and that the rest is rejected loudly instead of
being reinterpreted."""
import numpy as np
import pyarrow as pa
import pytest

from dpu_olap_b200 import ops


def test_sum_gpu_column_types():
    for dt in (np.uint32, np.int32, np.uint64, np.int64):
        s = ops.SumGpu(None, [np.arange(5, dtype=dt), pa.array(np.arange(3, dtype=dt))])
        assert s.dtype == np.dtype(dt)
    for bad in (np.float32, np.float64, np.uint16):
        with pytest.raises(TypeError):
            ops.SumGpu(None, [np.zeros(4, bad)])
    with pytest.raises(TypeError):
        ops.SumGpu(None, [np.zeros(4, np.uint32), np.zeros(4, np.int32)])      # mixed types
    with pytest.raises(TypeError):
        ops.SumGpu(None, [np.zeros(4, np.uint32), np.zeros(4, np.uint64)])     # mixed widths


def test_sum_gpu_keeps_arrow_offsets_and_bitmaps_of_64bit_columns():
    arr = pa.array([1, None, 3, 4, None, 6], type=pa.int64()).slice(1, 4)
    s = ops.SumGpu(None, [arr])
    col = s._ncols[0]
    assert col.dtype == np.dtype(np.int64) and col.values.dtype == np.dtype(np.uint64)
    assert col.values.tolist()[1:3] == [3, 4] and col.offset == 1 and col.valid is not None
    assert col.values.ctypes.data % 8 == 0


def test_take_gpu_column_types():
    idx = [np.array([0, 1], np.uint32)]
    assert ops.TakeGpu(None, [np.arange(4, dtype=np.uint32)], idx).dtype == np.dtype(np.uint32)
    assert ops.TakeGpu(None, [np.arange(4, dtype=np.uint64)], idx).dtype == np.dtype(np.uint64)
    assert ops.TakeGpu(None, [pa.array(np.arange(4, dtype=np.int64))], idx).dtype == np.dtype(np.int64)
    with pytest.raises(TypeError):
        ops.TakeGpu(None, [np.zeros(4, np.float64)], idx)
    with pytest.raises(TypeError):   # 64-bit values with nulls: not built
        ops.TakeGpu(None, [pa.array([1, None], type=pa.uint64())], idx)
    with pytest.raises(TypeError):   # 64-bit values, nullable indices
        ops.TakeGpu(None, [np.arange(4, dtype=np.uint64)], [pa.array([0, None], type=pa.uint32())])
    with pytest.raises(TypeError):   # indices are uint32
        ops.TakeGpu(None, [np.arange(4, dtype=np.uint64)], [np.array([0, 1], np.uint64)])
    with pytest.raises(ValueError):
        ops.TakeGpu(None, [np.zeros(4, np.uint32)], [])


def test_filter_gpu_column_types():
    for dt in (np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64):
        f = ops.FilterGpu(None, [np.zeros(4, dt), pa.array(np.zeros(3, dt))])
        assert f.dtype == np.dtype(dt)
    for bad in (np.uint16, np.int8, np.float16):
        with pytest.raises(TypeError):
            ops.FilterGpu(None, [np.zeros(4, bad)])
    with pytest.raises(TypeError):   # mixed widths
        ops.FilterGpu(None, [np.zeros(4, np.uint32), np.zeros(4, np.uint64)])
    with pytest.raises(TypeError):   # mixed types of one width
        ops.FilterGpu(None, [np.zeros(4, np.int64), np.zeros(4, np.float64)])
