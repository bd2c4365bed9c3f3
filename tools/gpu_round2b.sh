#!/bin/bash
# Round-2 final evidence run (under gpurun, ONE GPU): full GPU test suite, smoke, bench (own arm + reference arm), per-launch
# ncu metrics of one step of the three tuned BASELINE models, full-set captures of the fused bottleneck launches.
OUT=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/r02b_pytest_gpu_final.log; tail -3 $OUT/r02b_pytest_gpu_final.log
python __graft_entry__.py smoke > $OUT/r02b_smoke_final.log 2>&1; tail -2 $OUT/r02b_smoke_final.log
python bench.py --steps 30 --warmup 5 > $OUT/r02b_bench_final.json 2> $OUT/r02b_bench_final.err; tail -c 200 $OUT/r02b_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02b_bench_reference.json 2> $OUT/r02b_bench_reference.err; tail -c 300 $OUT/r02b_bench_reference.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for m in resnet50 vit_base efficientnet_b4; do
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/r02b_launch_metrics_$m.csv python tools/run_plan.py $m 3 > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bneck -c 3 -f -o $OUT/r02b_bneck python tools/run_plan.py resnet50 1 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep $OUT/r02b_launch_metrics_*.csv | tail -6
