#!/bin/bash
# Round-2 fifteenth GPU call: device input pipeline: bit-exactness tests against the PIL replay, timing.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_augment.py -m gpu -q -p no:cacheprovider --tb=short > $O/c15_augment.log 2>&1
echo "augment rc=$?" >> $O/c15_augment.log
tail -n 30 $O/c15_augment.log | cut -c1-400
timeout 300 python tests/diag/gpu_augment_bench.py 2>&1 | grep -h '^{\|Error\|error' | cut -c1-600 | tee $O/c15_augment_bench.log
