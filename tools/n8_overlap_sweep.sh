for k in 0 96 64; do
B2_PROBE_SCATTER_CTAS=$k python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --ops filter,join --no-cpu --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); j=d['ops']['join']; print('ctas',$k, j['ms_per_step'], j['self_check'], j['phases_ms_rank0_serialised'])"
done
