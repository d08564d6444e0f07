run() { # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-sampling 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$label', d['value'], d['ms_per_step'], d['timed_regions_ms_per_step']['value'])"
}
run default X=1
run ctas8_res8 NCCL_MAX_CTAS=8 BG_SM_RESERVE=8
run ctas4_res4 NCCL_MAX_CTAS=4 BG_SM_RESERVE=4
run ctas8 NCCL_MAX_CTAS=8
run ctas2_res2 NCCL_MAX_CTAS=2 BG_SM_RESERVE=2
