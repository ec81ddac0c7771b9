"""ctypes binding of libb200olap.so — the same C ABI (include/b200olap.h) the C++ host links.

There is deliberately NO fallback: if the CUDA library is missing or fails to load, importing
the operators raises, and every non-OK status raises B2Error.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
import os
# B200OLAP_LIB: the lab build (libb200olap_lab.so, -DB2_LAB) for tools/filter_lab.py; never set in tests or bench
LIB_PATH = Path(os.environ["B200OLAP_LIB"]).resolve() if os.environ.get("B200OLAP_LIB") else PKG_DIR / "libb200olap.so"
HEADER = PKG_DIR.parent / "include" / "b200olap.h"

B2_OK = 0
# enum b2_tunable
(TUNE_SCATTER_SECTORS_MIN_BITS, TUNE_SCATTER_PREFETCH, TUNE_SCATTER_SHAPE, TUNE_FILTER_VARIANT,
 TUNE_SCATTER_SECTOR_TILE, TUNE_JOIN_DIRECT_MIN_ROWS, TUNE_PEER_SCATTER_CTAS, TUNE_PEER_SCATTER_KERNEL,
 TUNE_FILTER64_KERNEL) = range(9)
STATUS_NAMES = {0: "B2_OK", 1: "B2_ERR_INVALID", 2: "B2_ERR_CUDA", 3: "B2_ERR_OOM",
                4: "B2_ERR_UNSUPPORTED", 5: "B2_ERR_WORKSPACE", 6: "B2_ERR_OVERFLOW"}


class B2Error(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)} {detail}".strip())


class Timings(C.Structure):
    _fields_ = [("copy_to_dev_ms", C.c_double), ("dev_work_ms", C.c_double),
                ("copy_from_dev_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_launches", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class JoinPhases(C.Structure):
    """b2_join_phases (include/b200olap.h): the join's phase timers, milliseconds."""
    _fields_ = [("partition_build_ms", C.c_double), ("partition_probe_ms", C.c_double),
                ("probe_ms", C.c_double), ("take_ms", C.c_double),
                ("intervals", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_vp, _i64, _u64, _u32, _int, _sz = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
_pp = C.POINTER(C.c_void_p)          # array of pointers
_pi64 = C.POINTER(C.c_int64)
_pu64 = C.POINTER(C.c_uint64)
_pu32 = C.POINTER(C.c_uint32)
_pt = C.POINTER(Timings)

# name -> (restype, argtypes); must cover every function include/b200olap.h declares
SIGNATURES = {
    "b2_version": (_int, []),
    "b2_strerror": (C.c_char_p, [_int]),
    "b2_device_count": (_int, [C.POINTER(C.c_int)]),
    "b2_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "b2_ctx_destroy": (_int, [_vp]),
    "b2_last_error": (C.c_char_p, [_vp]),
    "b2_launch_count": (_i64, [_vp]),
    "b2_ctx_device": (_int, [_vp]),
    "b2_ctx_sm_count": (_int, [_vp]),
    "b2_ctx_set_tunable": (_int, [_vp, _int, _int]),
    "b2_ctx_get_tunable": (_int, [_vp, _int, C.POINTER(C.c_int)]),
    "b2_gen_u32_dev": (_int, [_vp, _pu64, _pu32, _pu32, _i64, _i64, _vp, _vp]),
    "b2_iota_u32_dev": (_int, [_vp, _u64, _i64, _vp, _vp]),
    "b2_sum_u32_dev": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "b2_sum_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pu64, _pt]),
    "b2_filter_ws_bytes": (_sz, [_i64, _i64]),
    "b2_filter_lt_u32_dev": (_int, [_vp, _vp, _i64, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2_filter_ragged_ws_bytes": (_sz, [_pi64, _i64]),
    "b2_filter_lt_u32_ragged_dev": (_int, [_vp, _vp, _pi64, _vp, _i64, _u32, _vp, _vp, _vp, _vp,
                                           _vp, _sz, _vp]),
    "b2_filter_lt_u32_host": (_int, [_vp, _pp, _pi64, _i64, _u32, _pi64, _pu64, _pt]),
    "b2_filter_fetch_host": (_int, [_vp, _pp, _i64, _pt]),
    "b2_ctx_set_inputs_pinned": (_int, [_vp, _int]),
    "b2_sum_lt_u32_dev": (_int, [_vp, _vp, _i64, _u32, _vp, _vp, _vp]),
    "b2_host_alloc_pinned": (_int, [_sz, C.POINTER(C.c_void_p)]),
    "b2_host_free_pinned": (_int, [_vp]),
    "b2_host_register": (_int, [_vp, _sz]),
    "b2_host_unregister": (_int, [_vp]),
    "b2_shuffle_p2p_ws_bytes": (_sz, [_i64, _int]),
    "b2_shuffle_p2p_count_dev": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _sz, _vp]),
    "b2_shuffle_p2p_scatter_dev": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    "b2_shuffle_p2p_count_lt_dev": (_int, [_vp, _vp, _vp, _i64, _int, _int, _u32, _vp, _vp, _sz, _vp]),
    "b2_shuffle_p2p_scatter_lt_dev": (_int, [_vp, _vp, _vp, _i64, _int, _int, _u32, _vp, _vp, _vp, _sz, _vp]),
    "b2_shuffle_p2p_plan_dev": (_int, [_vp, _vp, _vp, _int, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b2_join_u32_phased_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _int, _int, _vp, _sz,
                                      _vp]),
    "b2_join_seg_cap_ws_bytes": (_sz, [_i64, _i64, _i64, _int, _int]),
    "b2_join_pairs_seg_cap_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _i64, _vp,
                                         _int, _vp, _vp, _sz, _vp]),
    "b2_join_pairs_seg_cap_phased_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _i64,
                                                _vp, _int, _vp, _int, _vp, _sz, _vp]),
    "b2_join_aggr_pairs_seg_cap_phased_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _int, _int, _u32, _vp,
                                                     _int, _vp, _int, _vp, _sz, _vp]),
    "b2_join_seg_ws_bytes": (_sz, [_i64, _i64, _int, _int]),
    "b2_join_pairs_seg_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _vp, _vp, _vp, _i64, _vp, _int,
                                     _vp, _sz, _vp]),
    "b2_filter_lt_u32_host_into": (_int, [_vp, _pp, _pi64, _i64, _u32, _vp, _i64, _pi64, _pu64, _pt]),
    "b2_take_u32_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "b2_take_64_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "b2_take_64_host": (_int, [_vp, _pp, _pi64, _pp, _pi64, _i64, _pp, _pt]),
    "b2_take_u32_ragged_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "b2_take_u32_host": (_int, [_vp, _pp, _pi64, _pp, _pi64, _i64, _pp, _pt]),
    "b2_wang_hash_u32": (_u32, [_u32]),
    "b2_partition_ws_bytes": (_sz, [_i64, _int]),
    "b2_partition_u32_dev": (_int, [_vp, _pp, _pp, _int, _i64, _int, _int, _vp, _vp, _sz, _vp]),
    "b2_partition_u32_host": (_int, [_vp, _pp, _pi64, _i64, _int, _int, _int, _pi64, _pt]),
    "b2_partition_fetch_host": (_int, [_vp, _pp, _int, _int, _pt]),
    "b2_join_ws_bytes": (_sz, [_i64, _i64]),
    "b2_join_min_ws_bytes": (_sz, [_i64, _i64]),
    "b2_join_ws_bytes_adjacent_outputs": (_sz, [_i64, _i64]),
    "b2_join_u32_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp,
                               _int, _vp, _sz, _vp]),
    "b2_join_pairs_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _int, _vp,
                                 _sz, _vp]),
    "b2_join_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pp, _pi64, _i64, _pu64, _pt]),
    "b2_join_fetch_host": (_int, [_vp, _vp, _vp, _vp, _i64, _pt]),
    "b2_join_dest_rank": (_int, [_u32, _int]),
    "b2_join_u32_nullable_host": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _pp, _pp, _pi64, _pi64, _i64, _pu64, _pt]),
    "b2_join_table_host": (_int, [_vp, _pp, _pi64, _i64, C.POINTER(C.c_int), _int, _pp, _pi64, _i64,
                                  C.POINTER(C.c_int), _int, _pu64, _pt]),
    "b2_join_table_fetch_host": (_int, [_vp, _pp, _int, _i64, _pt]),
    "b2_join_cols_u32_host": (_int, [_vp, _pp, _pi64, _i64, _int, _pp, _pi64, _i64, _int, _pu64, _pt]),
    "b2_join_cols_fetch_host": (_int, [_vp, _pp, _int, _i64, _pt]),
    "b2_join_aggr_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pp, _pi64, _i64, _int, _u32, _vp, _pt]),
    "b2_join_aggr_u32_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _u32, _vp, _int, _vp, _sz, _vp]),
    "b2_filter_lt_u32_nullable_dev": (_int, [_vp, _vp, _vp, _i64, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2_filter_lt_32_host_into": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _int, _u32, _vp, _i64, _pi64, _pu64, _pt]),
    "b2_filter_lt_32_ragged_dev": (_int, [_vp, _vp, _int, _u32, _vp, _pi64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2_join_trace": (_int, [_vp, _int]),
    "b2_join_last_phases": (_int, [_vp, C.POINTER(JoinPhases)]),
    "b2_filter_64_ws_bytes": (_sz, [_i64]),
    "b2_filter_lt_64_dev": (_int, [_vp, _vp, _int, _u64, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2_filter_lt_64_host_into": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _int, _u64, _vp, _i64, _pi64, _pu64, _pt]),
    "b2_filter_lt_32_dev": (_int, [_vp, _vp, _int, _u32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2_aggr_32_dev": (_int, [_vp, _vp, _int, _vp, _i64, _vp, _vp]),
    "b2_aggr_32_host": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _int, _vp, _pt]),
    "b2_aggr_u32_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "b2_aggr_64_dev": (_int, [_vp, _vp, _int, _vp, _i64, _vp, _vp]),
    "b2_aggr_64_host": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _int, _vp, _pt]),
    "b2_take_u32_nullable_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "b2_aggr_u32_host": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _vp, _pt]),
    "b2_filter_lt_u32_nullable_host_into": (_int, [_vp, _pp, _pp, _pi64, _pi64, _i64, _u32, _vp, _i64, _pi64,
                                                   _pu64, _pt]),
    "b2_take_u32_nullable_host": (_int, [_vp, _pp, _pp, _pi64, _pi64, _pp, _pp, _pi64, _pi64, _i64, _pp, _pp, _pt]),
    "b2_shuffle_ws_bytes": (_sz, [_i64, _int]),
    "b2_col_upload_host": (_int, [_vp, _pp, _pi64, _i64, C.POINTER(_vp)]),
    "b2_col_free": (_int, [_vp]),
    "b2_col_rows": (_i64, [_vp]),
    "b2_col_nbatches": (_i64, [_vp]),
    "b2_col_device_ptr": (_vp, [_vp]),
    "b2_col_batch_offsets": (_int, [_vp, _pi64, _i64]),
    "b2_col_download_host": (_int, [_vp, _vp, _pp, _i64]),
    "b2_col_export": (_int, [_vp, _vp]),
    "b2_col_import": (_int, [_vp, _vp, _pi64, _i64, C.POINTER(_vp)]),
    "b2_sum_u32_col": (_int, [_vp, _vp, _pu64]),
    "b2_filter_lt_u32_col": (_int, [_vp, _vp, _u32, C.POINTER(_vp)]),
    "b2_take_u32_col": (_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "b2_join_u32_col": (_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "b2_set_create": (_int, [C.POINTER(C.c_int), _int, C.POINTER(_vp)]),
    "b2_set_destroy": (_int, [_vp]),
    "b2_set_size": (_int, [_vp]),
    "b2_set_ctx": (_vp, [_vp, _int]),
    "b2_set_peer_access": (_int, [_vp]),
    "b2_set_last_error": (C.c_char_p, [_vp]),
    "b2_set_launch_count": (_i64, [_vp]),
    "b2_set_set_inputs_pinned": (_int, [_vp, _int]),
    "b2_set_sum_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pu64, _pt]),
    "b2_set_filter_lt_u32_host": (_int, [_vp, _pp, _pi64, _i64, _u32, _pi64, _pu64, _pt]),
    "b2_set_filter_fetch_host": (_int, [_vp, _pp, _i64, _pt]),
    "b2_set_take_u32_host": (_int, [_vp, _pp, _pi64, _pp, _pi64, _i64, _pp, _pt]),
    "b2_set_join_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pp, _pi64, _i64, _pu64, _pt]),
    "b2_set_join_aggr_u32_host": (_int, [_vp, _pp, _pi64, _i64, _pp, _pi64, _i64, _int, _u32, _vp, _pt]),
    "b2_set_join_fetch_host": (_int, [_vp, _vp, _vp, _vp, _i64, _pt]),
    "b2_shuffle_partition_u32_dev": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
}


def header_functions() -> list[str]:
    """Names of every function declared in include/b200olap.h (used by the export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Load libb200olap.so (built in-tree by dpu_olap_b200/build.py). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dpu_olap_b200.build` "
            "(there is no CPU fallback for the operator path)")
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def check(status: int, where: str, ctx: C.c_void_p | None = None) -> None:
    if status != B2_OK:
        detail = ""
        if ctx:
            detail = lib().b2_last_error(ctx).decode(errors="replace")
        raise B2Error(status, where, detail)
