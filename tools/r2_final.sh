#!/bin/bash
# Final validation + profile capture of round 2 on one B200 (run through gpurun).  A/B of the pair loop first (variants/pl0.so =
# the previous traverse.cu / forces.cu), then the GPU tests and smoke() on the default build, the ncu capture of the dominant
# kernel (-> profiles/r2_traffic.json with the hash of the traverse.cu it ran), the default bench line, the ncu launch list.
mkdir -p gpurun_out
V=$PWD/naivedynamics.jl_b200/variants
for k in 1; do
  [ -f $V/pl0.so ] && NAIVEB200_LIB=$V/pl0.so NB200_PRESTEPS=300 timeout 100 python tools/stage_bench.py c3 100 2>&1 | tail -1
  NB200_PRESTEPS=300 timeout 100 python tools/stage_bench.py c3 100 2>&1 | tail -1
done
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
NB200_NO_GRAPH=1 NB200_PRESTEPS=600 timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 610 -c 1 -f \
    -o gpurun_out/r2_traverse_final python tools/stage_bench.py c3 30 2>&1 | tail -2
python tools/traffic_from_ncu.py gpurun_out/r2_traverse_final.ncu-rep profiles/r2_traffic.json > /dev/null && cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 300 gpurun_out/r2_bench_1gpu.err
NB200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 20 --warmup 3 --melt 60 --cpu-budget 1 > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_1gpu.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['traffic'], d['e2e']['value'], {k:v.get('value', v.get('searches_per_s')) for k,v in d['variants'].items()})"
