#!/bin/bash
# The round-2 evidence run on one B200 (through gpurun): optional A/B of library variants (every naivedynamics.jl_b200/variants/*.so
# against the default build, tools/variants.sh builds them), smoke(), the GPU tests, then — in this order, so that the bench line
# can carry the traffic of the build it ran — one `ncu --set full` capture of the dominant kernel -> profiles/r2_traffic.json
# (tools/traffic_from_ncu.py: DRAM bytes + the sha1 of the traverse.cu they were measured on), the default bench line, and the ncu
# launch list of a short bench run.  Outputs under gpurun_out/; copy what is to be judged into profiles/.
# usage: bash tools/r2_profile.sh [notests]
mkdir -p gpurun_out
sb() { NB200_PRESTEPS=300 timeout 100 python tools/stage_bench.py c3 100 2>&1 | tail -1; }
for v in naivedynamics.jl_b200/variants/*.so; do [ -f "$v" ] && NAIVEB200_LIB=$PWD/$v sb; done
sb
NB200_FUSED=0 sb
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
[ "$1" = notests ] || timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
NB200_NO_GRAPH=1 NB200_PRESTEPS=600 timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 610 -c 1 -f \
    -o gpurun_out/r2_traverse_final python tools/stage_bench.py c3 30 2>&1 | tail -1
python tools/traffic_from_ncu.py gpurun_out/r2_traverse_final.ncu-rep profiles/r2_traffic.json > /dev/null && cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 300 gpurun_out/r2_bench_1gpu.err
NB200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 20 --warmup 3 --melt 60 --cpu-budget 1 > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_1gpu.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['traffic'], d['e2e']['value'], d['energy'], {k:v.get('value', v.get('searches_per_s')) for k,v in d['variants'].items()})"
