#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: the fused bottleneck (CTA pair, tensor-memory operand, in-place
# residual), the sixteen-warp epilogue + gathering producers (forced with EQXV_EPI_SUB=4 on small shapes), the K tail
# shift, the gated narrow-K path and the first-layer kernel with the max-pool epilogue. Logs -> gpurun_out/, digests ->
# profiles/r02_sanitizer_*_late.log.
set -u
OUT=gpurun_out
SEL='bottleneck64 and (2x56x56 or 3x24x40 or 2x30x23 or 1x56x56 or refuses) or k_tail_shift or se_gate and (3-100-40 or 3-1000-56 or 5-196-672) or stem_with_fused_maxpool and 2-64-64'
SEL16='test_gemm and (512x24x144 or 512x64x24 or 128x64x64 or 1000x128x192 or 300x272x1632) or conv2d_igemm and (3x17x13 or 2x30x23 or 4x56x56x64x256)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/sanitizer_memcheck_late.log 2>&1; echo "memcheck rc=$?" >> $OUT/sanitizer_memcheck_late.log
EQXV_EPI_SUB=4 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL16" > $OUT/sanitizer_memcheck_epi16.log 2>&1; echo "memcheck(epi16 forced) rc=$?" >> $OUT/sanitizer_memcheck_epi16.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/sanitizer_synccheck_late.log 2>&1; echo "synccheck rc=$?" >> $OUT/sanitizer_synccheck_late.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/sanitizer_racecheck_late.log 2>&1; echo "racecheck rc=$?" >> $OUT/sanitizer_racecheck_late.log
for f in memcheck_late memcheck_epi16 synccheck_late racecheck_late; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Error|hazard" $OUT/sanitizer_$f.log | tail -6; done
