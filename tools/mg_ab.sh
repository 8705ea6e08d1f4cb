# A/B runs of the 2-GPU slab bench (tuning aid): each line = one configuration (environment variables), ms/step, rank-0 stage split
run() { env "$@" NB200_NO_VARIANTS=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 --melt 300 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['ms_per_step'],4), d['roofline']['stage_ms_per_step_rank0'], d['parity']['count_match'] and d['parity']['xor_match'])
"; }
for cfg in "$@"; do run $cfg; done
