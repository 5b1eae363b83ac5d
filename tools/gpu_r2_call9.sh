#!/bin/bash
# Round-2 ninth GPU call: first GPU run of the DeepLabV3+/WRN-38 path: op tests, model tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_f2_ops.py -m gpu -q --tb=short -p no:cacheprovider > $O/c9_f2_ops.log 2>&1
echo "f2 ops rc=$?" >> $O/c9_f2_ops.log
timeout 400 python -m pytest tests/test_gpu_deepv3.py -m gpu -q --tb=short -p no:cacheprovider > $O/c9_deepv3.log 2>&1
echo "deepv3 rc=$?" >> $O/c9_deepv3.log
tail -n 40 $O/c9_f2_ops.log | cut -c1-300
tail -n 40 $O/c9_deepv3.log | cut -c1-300
