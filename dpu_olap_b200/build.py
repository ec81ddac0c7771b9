"""Build recipe for libb200olap.so (hand-written sm_100a CUDA, C ABI in include/b200olap.h).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb200olap.so"

SOURCES = ["ctx.cu", "sum.cu", "filter.cu", "take.cu", "gen.cu", "scan.cu", "partition.cu",
           "join.cu", "nullable.cu", "api_host.cu", "api_host_join.cu", "set.cu", "col.cu", "filter64.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-cudart", "shared",
    "-Xptxas", "-v",
]
OBJ_DIR = PKG_DIR / "build"


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _headers_digest(extra_flags: list[str]) -> "hashlib._Hash":
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [REPO / "include" / "b200olap.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + extra_flags).encode())
    return h


def build(force: bool = False, verbose: bool = False, lab: bool = False) -> Path:
    """Compile every .cu in csrc/ (one object per file, in parallel, only what changed) and link
    dpu_olap_b200/libb200olap.so. lab=True builds libb200olap_lab.so with -DB2_LAB: the switches of
    tools/filter_lab.py that make kernels compute wrong results on purpose exist only there."""
    from concurrent.futures import ThreadPoolExecutor
    extra = ["-DB2_LAB"] if lab else []
    lib_path = PKG_DIR / ("libb200olap_lab.so" if lab else "libb200olap.so")
    obj_dir = OBJ_DIR / ("lab" if lab else "product")
    obj_dir.mkdir(parents=True, exist_ok=True)
    hdr = _headers_digest(extra)
    nvcc = nvcc_path()
    logs: dict[str, str] = {}

    def compile_one(src: Path) -> tuple[Path, bool]:
        h = hdr.copy()
        h.update(src.read_bytes())
        digest = h.hexdigest()
        obj, stamp = obj_dir / (src.stem + ".o"), obj_dir / (src.stem + ".stamp")
        if not force and obj.exists() and stamp.exists() and stamp.read_text().strip() == digest:
            logs[src.name] = (obj_dir / (src.stem + ".log")).read_text() if (obj_dir / (src.stem + ".log")).exists() else ""
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(REPO / "include"), "-c", "-o", str(obj), str(src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log = " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        logs[src.name] = log
        (obj_dir / (src.stem + ".log")).write_text(log)
        if proc.returncode != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"nvcc failed on {src.name}")
        stamp.write_text(digest)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, _sources()))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not lib_path.exists() or force:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "shared",
               "-o", str(lib_path), *[str(o) for o in objs]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"linking {lib_path.name} failed")
    if not lab:
        (PKG_DIR / "build.log").write_text("\n".join(logs[k] for k in sorted(logs)))
    if verbose:
        sys.stderr.write("\n".join(logs[k] for k in sorted(logs)))
    return lib_path


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, lab="--lab" in sys.argv)
    print(p)
