"""The C++ host layer (dpu_olap_b200/host): builds against Arrow C++ 24 from the pyarrow wheel.

CPU: the generator restatement reproduces the golden fingerprints and the Native (Acero) classes
return the reference tests' known answers. GPU: the reference's GoogleTests restated over the
*Gpu operator classes (host_test) and the benchmark driver's gbench-shaped JSON (host_bench).
"""
import json
import os
import subprocess

import pytest


@pytest.fixture(scope="module")
def host_bins(built_lib):
    from dpu_olap_b200.host import build_host
    return {p.name: p for p in build_host.build()}


def test_cpp_generator_and_native_known_answers(host_bins):
    r = subprocess.run([str(host_bins["host_test"]), "--cpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_bench_native_json_shape(host_bins):
    env = dict(os.environ, SF="1", MAX_THREADS="2")
    r = subprocess.run([str(host_bins["host_bench"]), "--benchmark_filter=BM_FilterNative", "--iterations=1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    assert d["context"]["SF"] == "1"
    (b,) = d["benchmarks"]
    # scripts/parse_results.py:25-35 splits the name on "/" and ":" and takes part [1] as the operator
    parts = b["name"].split("/")
    assert parts[1] == "BM_FilterNative" and parts[2] == "Batches:128" and not b["error_occurred"]
    assert b["items_per_second"] > 0


@pytest.mark.gpu
def test_cpp_reference_tests_on_gpu(host_bins):
    r = subprocess.run([str(host_bins["host_test"])], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 of 18 cases failed" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("nr_gpus", [2, 8])
def test_cpp_reference_tests_on_a_gpu_set(host_bins, nr_gpus):
    """The same GoogleTest restatements with GpuSet::allocate(NR_GPUS): every operator class shards
    over the whole set behind the C ABI (JoinTest.LargeTest runs the fused peer-store shuffle)."""
    import torch
    if torch.cuda.device_count() < nr_gpus:
        pytest.skip(f"needs {nr_gpus} GPUs")
    env = dict(os.environ, NR_GPUS=str(nr_gpus))
    r = subprocess.run([str(host_bins["host_test"])], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 of 18 cases failed" in r.stdout


@pytest.mark.gpu
def test_cpp_bench_gpu_cases(host_bins):
    env = dict(os.environ, SF="2")
    r = subprocess.run([str(host_bins["host_bench"]), "--iterations=1"], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    names = {b["name"].split("/")[1] for b in d["benchmarks"]}
    assert {"BM_FilterGpu", "BM_SumGpu", "BM_TakeGpu", "BM_JoinGpu", "BM_PartitionGpu", "BM_FilterNative",
            "BM_JoinNative"} <= names
    assert not any(b["error_occurred"] for b in d["benchmarks"])
    gpu = [b for b in d["benchmarks"] if b["name"].split("/")[1].endswith("Gpu")]
    assert all("dpu-work" in b and "copy-to-dpu" in b for b in gpu)


def test_reference_parse_results_reads_our_benchmark_json(host_bins, tmp_path):
    """The UNMODIFIED reference script scripts/parse_results.py turns the driver's JSON into the
    CSV its plots are made from (operator column from name part [1], one column per Arg)."""
    import csv
    import sys
    from pathlib import Path
    script = Path("/root/reference/scripts/parse_results.py")
    if not script.exists():
        pytest.skip("the reference tree is not available here")
    env = dict(os.environ, SF="1", MAX_THREADS="2")
    out = tmp_path / "filter_native_1.json"
    r = subprocess.run([str(host_bins["host_bench"]), "--benchmark_filter=BM_FilterNative",
                        f"--benchmark_out={out}", "--benchmark_out_format=json", "--benchmark_repetitions=1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and out.exists(), r.stderr
    r = subprocess.run([sys.executable, str(script), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = list(csv.DictReader(open(tmp_path / "filter_native_1.csv")))
    assert len(rows) == 1 and rows[0]["operator"] == "BM_FilterNative"
    assert rows[0]["Batches"] == "128" and rows[0]["BatchSize"] == "65536" and rows[0]["Threads"] == "2"
    assert float(rows[0]["items_per_second"]) > 0
