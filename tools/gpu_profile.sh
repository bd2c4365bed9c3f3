#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per launch for one bench step,
# (2) one full-set capture of the dominant kernel (igemm) on three representative launches.
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 60 -c 12 \
    -o gpurun_out/${TAG}_igemm python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline \
    > gpurun_out/${TAG}_igemm_bench.log 2>&1
ls -la gpurun_out | tail -8
