#!/bin/bash
# Final 2-GPU run of round 2 (through gpurun --gpus 2): the default multi-GPU bench line (headline + the 8M strong-scaling
# variant, in-run parity), then the same headline with NB200_CARVEOUT=100 (one shared-memory carve-out for every kernel of the
# slab step) for comparison.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -c 300 gpurun_out/r2_bench_2gpu.err
NB200_CARVEOUT=100 NB200_NO_VARIANTS=1 timeout 200 $TR bench.py --gpus 2 > gpurun_out/r2_bench_2gpu_carveout100.json 2> gpurun_out/r2_bench_2gpu_carveout100.err
python - <<'P'
import json
for f in ("r2_bench_2gpu", "r2_bench_2gpu_carveout100"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["ms_per_step"], 4), "%.4g" % d["value"], d["roofline"]["stage_ms_per_step_rank0"], d["parity"]["count_match"] and d["parity"]["xor_match"],
              {k: (v.get("ms_per_step"), (v.get("parity") or {}).get("xor_match"), v.get("error")) for k, v in d["variants"].items() if isinstance(v, dict)}, d["variants"].get("error"))
    except Exception as e:
        print(f, "FAILED", e)
P
