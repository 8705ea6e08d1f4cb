#!/bin/bash
# Round-2 validation + profile capture on one B200 (run through gpurun): the GPU tests, smoke(), the default bench line,
# the ncu launch list of a short bench run and one `ncu --set full` capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 300 gpurun_out/r2_bench_1gpu.err
NB200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 20 --warmup 3 --melt 60 --cpu-budget 1 > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
NB200_NO_GRAPH=1 NB200_PRESTEPS=600 timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 610 -c 1 -f \
    -o gpurun_out/r2_traverse_fused python tools/stage_bench.py c3 30 2>&1 | tail -3
ls -la gpurun_out/r2_*
