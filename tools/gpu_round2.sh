#!/bin/bash
# Round-2 evidence run (under gpurun, ONE GPU): full GPU test suite, smoke, bench, per-launch ncu metrics of one step of
# the three tuned BASELINE models (time, DRAM bytes, tensor-pipe activity), full-set captures of the kernels VERDICT r1
# asked for. Everything lands in gpurun_out/; tools/launch_summary.py / tools/ncu_digest.py turn it into profiles/.
OUT=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/r02_pytest_gpu_final.log; tail -3 $OUT/r02_pytest_gpu_final.log
python __graft_entry__.py smoke > $OUT/r02_smoke_final.log 2>&1; tail -2 $OUT/r02_smoke_final.log
python bench.py --steps 30 --warmup 5 > $OUT/r02_bench_final.json 2> $OUT/r02_bench_final.err; tail -c 200 $OUT/r02_bench_final.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for m in resnet50 vit_base efficientnet_b4; do
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/r02_launch_metrics_$m.csv python tools/run_plan.py $m 3 > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 14 -c 1 -f -o $OUT/r02_attention_pp python tools/run_plan.py vit_base 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 60 -c 4 -f -o $OUT/r02_pair_vit python tools/run_plan.py vit_base 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 2 -c 1 -f -o $OUT/r02_layernorm python tools/run_plan.py vit_base 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eltwise -s 40 -c 1 -f -o $OUT/r02_eltwise_b4 python tools/run_plan.py efficientnet_b4 2 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep $OUT/r02_launch_metrics_*.csv
