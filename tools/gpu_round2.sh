#!/bin/bash
# tests + bench + A/B switches + time-only launch list of one model (cheap ncu pass)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py ${BENCH_FLAGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for sw in ${AB_SWITCHES}; do
  env $sw timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_${sw%%=*}.json 2>&1
  echo "== $sw"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${sw%%=*}.json"))
    print(d["value"], d["ms_per_step"], {k:(v["value"],v["ms_per_step"]) for k,v in d["secondary"].items()})
except Exception as e:
    print("failed", e)
PY
done
for model in ${NCU_MODELS}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${model}.csv \
      python bench.py --model $model --steps 1 --warmup 3 --no-secondary --no-cpu-baseline \
      > gpurun_out/${TAG}_launches_${model}.log 2>&1
done
ls -la gpurun_out | tail -8
