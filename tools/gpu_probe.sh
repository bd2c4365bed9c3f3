#!/bin/bash
# Runs every probe group in its own process with a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/probe_*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/probe_gpu.log 2>&1
for g in ${@:-gemm conv conv_s2 stem pointwise attention}; do
  timeout 300 python tools/probe_kernels.py $g > gpurun_out/probe_$g.log 2>&1
  echo "$g exit $?" >> gpurun_out/probe_summary.log
done
tail -n 60 gpurun_out/probe_*.log
