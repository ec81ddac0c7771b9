"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a without a GPU, loads,
and exports every symbol include/b200olap.h declares; the ctypes table covers the header; no
compute call is made here."""
import ctypes as C
import re
import subprocess

import pytest

from dpu_olap_b200 import _lib


def test_library_builds_and_loads(built_lib):
    assert built_lib.exists()
    handle = _lib.lib()
    assert handle.b2_version() == 100
    assert handle.b2_strerror(0) == b"ok"
    assert handle.b2_strerror(5) == b"workspace too small"


def test_exports_every_declared_symbol(built_lib):
    declared = _lib.header_functions()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r"\bT (b2_[a-z0-9_]+)", out))
    assert not [f for f in declared if f not in exported]
    # nothing but the C ABI leaks b2_* names with C linkage that the header does not declare
    assert exported - set(declared) <= set()


def test_ctypes_table_matches_header(built_lib):
    assert sorted(_lib.SIGNATURES) == _lib.header_functions()


def test_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "--list-elf", str(built_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_side_pure_functions(built_lib):
    h = _lib.lib()
    # reference hash (dpu/shared/kernels/partition.c:20-28); SURVEY.md §8(c) fingerprints
    assert h.b2_wang_hash_u32(0) == 0x4636B9C9
    assert h.b2_wang_hash_u32(2) == 0xFF4D1170
    # partition_test.cc:28-36: {0,2,3,8} with 2 partitions -> 3 / 1
    assert [h.b2_join_dest_rank(k, 2) for k in (0, 2, 3, 8)].count(0) == 3
    assert h.b2_filter_ws_bytes(4, 65536) >= 64 + 4 * 8 * 8
    assert h.b2_join_min_ws_bytes(1 << 22, 1 << 22) < h.b2_join_ws_bytes(1 << 22, 1 << 22)


def test_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: without a usable GPU the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dpu_olap_b200._lib import B2Error
    from dpu_olap_b200.ops import Context
    with pytest.raises(B2Error):
        Context(0)


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libb200olap.so")
    with pytest.raises(ImportError):
        _lib.lib()
