#!/bin/bash
# compute-sanitizer over the kernel-level parity tests ("race detection"): memcheck on every kernel family,
# racecheck + synccheck on the kernels that hand-roll mbarrier / TMEM protocols (igemm, pair, halo, attention) at
# reduced shapes. Run under gpurun; logs land in gpurun_out/ and are summarised into profiles/.
set -u
OUT=gpurun_out
SEL_ALL='conv2d_igemm and (2x56x56x64x64x3 or 4x56x56x64x256 or 5x7x7x512 or 3x17x13 or 2x16x16x64x64x3x1x12 or 2x56x56x256x512 or 2x224x224x8x48 or 1x56x56x128x320 or 2x30x23) or gemm and (128x64x64 or 1000x128x192 or 300x272x1632 or 7x448x8 or 256x1000x2048) or test_stem or layernorm or attention or depthwise or eltwise or global_avgpool or window_attention or vit_glue or pooling_and_layout'
SEL_RACE='conv2d_igemm and (2x56x56x64x64x3 or 4x56x56x64x256 or 5x7x7x512 or 3x17x13) or gemm and (128x64x64 or 1000x128x192 or 300x272x1632) or test_attention and 197 or depthwise_with_fused_squeeze and (40 or 72)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py tests/test_input_edge.py -m gpu -q -x -k "$SEL_ALL or u8_kernels" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL_RACE" > $OUT/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> $OUT/sanitizer_synccheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL_RACE" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck.log
for f in memcheck synccheck racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Error" $OUT/sanitizer_$f.log | tail -6; done
