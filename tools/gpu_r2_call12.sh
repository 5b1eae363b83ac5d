#!/bin/bash
# Round-2 twelfth GPU call: deferred BatchNorm finalisation (op tests, A/B on the step), remaining launch-width knobs,
# then the whole -m gpu suite of the branch (f2 tests with the noise-floor criterion included).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider -k "deferred or batchnorm_train or in_launch" > $O/c12_ops.log 2>&1
echo "ops rc=$?" >> $O/c12_ops.log
tail -n 5 $O/c12_ops.log | cut -c1-300
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
export B200SEG_TIME_ONLY=1
for cfg in "B200SEG_BN_CELLS=1" "B200SEG_BN_CELLS=0" "B200SEG_BN_CELLS=1 B200SEG_EW_ITEMS=16" "B200SEG_BN_CELLS=1 B200SEG_RED_ITEMS=8" \
           "B200SEG_BN_CELLS=1 B200SEG_RS_ITEMS=4" "B200SEG_BN_CELLS=1 B200SEG_EW_ITEMS=4" "B200SEG_BN_CELLS=1"; do
  echo "== $cfg" >> $O/c12_ab.log
  env $cfg timeout 200 $B 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300 >> $O/c12_ab.log
done
cat $O/c12_ab.log
unset B200SEG_TIME_ONLY
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 > $O/c12_suite.log 2>&1
echo "suite rc=$?" >> $O/c12_suite.log
tail -n 25 $O/c12_suite.log | cut -c1-300
