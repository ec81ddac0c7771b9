#!/usr/bin/env bash
# Sweep runner in the shape of the reference's scripts/run-cpp-baselines.sh / run-upmem-2048.sh:
# one JSON report per operator and scale factor from the C++ benchmark driver (BM_*Gpu beside
# BM_*Native on the same generator(42) inputs), then scripts/parse_results.py-compatible CSVs, and
# the device-resident scaling line of bench.py for 1/2/4/8 GPUs.
#
#   scripts/run-gpu-sweep.sh [max_sf=64] [gpus="1 2 4 8"]
set -u
MAX_SF=${1:-64}
GPUS=${2:-"1 2 4 8"}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
BIN=$ROOT/dpu_olap_b200/host/_build/host_bench
mkdir -p "$ROOT/reports"
python -m dpu_olap_b200.build >/dev/null && python -m dpu_olap_b200.host.build_host >/dev/null || exit 1

for op in Filter Sum Take Join; do
  for kind in Native Gpu; do
    sf=1
    while [ "$sf" -le "$MAX_SF" ]; do
      SF=$sf "$BIN" --benchmark_filter=BM_${op}${kind} --benchmark_repetitions=3 \
        --benchmark_out="$ROOT/reports/$(echo $op | tr A-Z a-z)_$(echo $kind | tr A-Z a-z)_$sf.json" --benchmark_out_format=json
      sf=$((sf * 2))
    done
  done
done

# the reference's own converter works on these files (tests/test_host_cpp.py checks that)
if [ -f /root/reference/scripts/parse_results.py ]; then
  python /root/reference/scripts/parse_results.py "$ROOT/reports"
fi

for n in $GPUS; do
  if [ "$n" -eq 1 ]; then
    python "$ROOT/bench.py" --gpus 1 > "$ROOT/reports/bench_n1.json"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port $((29600 + n)) "$ROOT/bench.py" --gpus "$n" > "$ROOT/reports/bench_n$n.json"
  fi
done
