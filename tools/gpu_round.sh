#!/bin/bash
# One GPU session: tests, bench (+A/B switches), per-launch ncu metric passes. Logs under gpurun_out/<tag>_*.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --all-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for sw in ${AB_SWITCHES}; do
  env $sw timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_${sw%%=*}.json 2>&1
  echo "== $sw"; cut -c1-400 gpurun_out/${TAG}_bench_${sw%%=*}.json
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed"
for model in ${NCU_MODELS:-resnet50 vit_base efficientnet_b4}; do
  timeout 400 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${model}.csv \
      python bench.py --model $model --steps 1 --warmup 3 --no-secondary --no-cpu-baseline \
      > gpurun_out/${TAG}_launches_${model}.log 2>&1
done
ls -la gpurun_out | tail -12
