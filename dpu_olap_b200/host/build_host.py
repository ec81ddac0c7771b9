"""Builds the C++ host layer with g++ against the Arrow C++ that ships inside the pyarrow wheel
(headers + libarrow*.so.2400) and against libb200olap.so. No CMake: five translation units."""
from __future__ import annotations

import hashlib
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
REPO = PKG.parent
OUT = HERE / "_build"
COMMON = ["operators.cc", "native.cc", "generator.cc"]
PROGRAMS = {"host_test": "host_test.cc", "host_bench": "host_bench.cc"}


def arrow_paths() -> tuple[Path, list[str]]:
    import pyarrow
    root = Path(pyarrow.__file__).resolve().parent
    libs = []
    for stem in ("arrow", "arrow_compute", "arrow_acero"):
        cands = sorted(root.glob(f"lib{stem}.so.*"))
        if not cands:
            raise RuntimeError(f"lib{stem}.so not found in {root}")
        libs.append(cands[0].name)
    return root, libs


def _digest(srcs: list[Path]) -> str:
    h = hashlib.sha256()
    for f in sorted(srcs + list(HERE.glob("*.h")) + [REPO / "include" / "b200olap.h"]):
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False) -> list[Path]:
    root, libs = arrow_paths()
    OUT.mkdir(exist_ok=True)
    built = []
    for prog, main in PROGRAMS.items():
        srcs = [HERE / s for s in COMMON + [main]]
        exe = OUT / prog
        stamp = OUT / f".{prog}.stamp"
        dg = _digest(srcs)
        if not force and exe.exists() and stamp.exists() and stamp.read_text() == dg:
            built.append(exe)
            continue
        cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wno-deprecated-declarations",
               "-I", str(root / "include"), "-I", str(REPO / "include"), "-I", str(HERE),
               *[str(s) for s in srcs], "-o", str(exe),
               "-L", str(root), *[f"-l:{lib}" for lib in libs],
               "-L", str(PKG), "-l:libb200olap.so",
               f"-Wl,-rpath,{root}", "-Wl,-rpath,$ORIGIN/../..", "-Wl,--allow-shlib-undefined", "-pthread"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        (OUT / f"{prog}.log").write_text(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"g++ failed building {prog}")
        stamp.write_text(dg)
        built.append(exe)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print(p)
