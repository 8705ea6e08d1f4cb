#!/bin/bash
# A/B of the fused traversal's blocks per SM (variants/mb12.so = 12, default = 14, variants/mb15.so = 15) on one B200, the stage
# split of the un-fused configuration, smoke(); if the default build is the faster one: ncu capture of the dominant kernel
# (-> profiles/r2_traffic.json), the default bench line and the ncu launch list, as tools/r2_final.sh.
mkdir -p gpurun_out
V=$PWD/naivedynamics.jl_b200/variants
sb() { NB200_PRESTEPS=300 timeout 100 python tools/stage_bench.py c3 100 2>&1 | tail -1; }
NAIVEB200_LIB=$V/mb12.so sb | tee gpurun_out/ab_mb12.txt
sb | tee gpurun_out/ab_default.txt
NAIVEB200_LIB=$V/mb15.so sb | tee gpurun_out/ab_mb15.txt
NB200_FUSED=0 sb | tee gpurun_out/ab_unfused.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
python - <<'P' || exit 0
import re, sys
t = lambda f: float(re.search(r"no events ([0-9.]+)", open(f"gpurun_out/{f}.txt").read()).group(1))
a, b = t("ab_mb12"), t("ab_default")
print("12 blocks/SM %.4f ms/step, 14 blocks/SM %.4f" % (a, b))
sys.exit(0 if b < a - 0.002 else 1)
P
NB200_NO_GRAPH=1 NB200_PRESTEPS=600 timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 610 -c 1 -f \
    -o gpurun_out/r2b_traverse_final python tools/stage_bench.py c3 30 2>&1 | tail -1
python tools/traffic_from_ncu.py gpurun_out/r2b_traverse_final.ncu-rep profiles/r2_traffic.json > /dev/null && cp profiles/r2_traffic.json gpurun_out/r2b_traffic.json
timeout 400 python bench.py > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; tail -c 300 gpurun_out/r2b_bench_1gpu.err
NB200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 20 --warmup 3 --melt 60 --cpu-budget 1 > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_1gpu.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['traffic'], d['e2e']['value'], d['energy'], {k:v.get('value', v.get('searches_per_s')) for k,v in d['variants'].items()})"
