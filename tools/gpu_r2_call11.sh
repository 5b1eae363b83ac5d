#!/bin/bash
# Round-2 eleventh GPU call: (1) DeepLabV3+ gradient diagnostic in four conditionings, (2) coordinate search over the
# launch-width knobs on the two-scale train step, (3) op / block tests under the narrowest setting tried (grid-independence).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tests/diag/gpu_deepv3_diag.py > $O/c11_deepv3_diag.log 2>&1
echo "diag rc=$?" >> $O/c11_deepv3_diag.log
grep '^\[' $O/c11_deepv3_diag.log | cut -c1-400
timeout 700 python tools/gpu_tune_sweep.py --budget-s 400 > $O/c11_sweep.log 2>&1
echo "sweep rc=$?" >> $O/c11_sweep.log
cat $O/c11_sweep.log | cut -c1-200
B200SEG_WGRAD_MIN_CLK=150000 B200SEG_WGRAD_STREAMS=3 B200SEG_CONV_MIN_CLK=40000 B200SEG_EW_ITEMS=8 B200SEG_EW_CTAS_PER_SM=2 \
  B200SEG_RED_ITEMS=4 B200SEG_RED_CTAS_PER_SM=1 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py \
  -m gpu -x -q -p no:cacheprovider > $O/c11_narrow_tests.log 2>&1
echo "narrow tests rc=$?" >> $O/c11_narrow_tests.log
tail -n 6 $O/c11_narrow_tests.log | cut -c1-300
