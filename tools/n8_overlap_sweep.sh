#!/bin/bash
# Sweep of the overlapped sharded join on 8 GPUs: probe shares x CTA budget of the probe side's scatter.
#   gpurun --gpus 8 -- bash tools/n8_overlap_sweep.sh "2:96 2:120 3:96"
for cfg in ${1:-"1:0 1:96 2:96"}; do
  s=${cfg%%:*}; k=${cfg##*:}
  B2_PROBE_SHARES=$s B2_PROBE_SCATTER_CTAS=$k python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --ops filter,join --no-cpu --no-e2e --steps 10 2>/dev/null |
    python -c "
import json,sys
d=json.loads(sys.stdin.read()); j=d['ops']['join']
print('shares', $s, 'ctas', $k, round(j['ms_per_step'], 3), j['self_check'], j['phases_ms_rank0_serialised'])"
done
