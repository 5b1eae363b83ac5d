#!/bin/bash
# Round-2 fourteenth GPU call: (1) late programmatic-launch trigger (build variant libb200seg_late.so) A/B on the step;
# (2) evaluation with conv + BN + residual + ReLU in one launch: cfg5 timing A/B, then the whole -m gpu suite with it on.
set -u
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
export B200SEG_TIME_ONLY=1
for cfg in "B200SEG_LIB_VARIANT=" "B200SEG_LIB_VARIANT=late" "B200SEG_LIB_VARIANT=" "B200SEG_LIB_VARIANT=late"; do
  echo "== $cfg"
  env $cfg timeout 200 $B 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300
done | tee $O/c14_pdl_ab.log
unset B200SEG_TIME_ONLY
B200SEG_LIB_VARIANT=late timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py tests/test_gpu_model.py -m gpu -x -q -p no:cacheprovider > $O/c14_late_tests.log 2>&1
echo "late tests rc=$?" >> $O/c14_late_tests.log
tail -n 4 $O/c14_late_tests.log | cut -c1-300
for f in 0 1; do B200SEG_EVAL_FUSED=$f timeout 300 python tools/gpu_eval_bench.py 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300; done | tee $O/c14_eval_bench.log
B200SEG_EVAL_FUSED=1 timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/c14_suite.log 2>&1
echo "suite rc=$?" >> $O/c14_suite.log
tail -n 8 $O/c14_suite.log | cut -c1-300
