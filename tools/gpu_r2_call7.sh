#!/bin/bash
# Round-2 seventh GPU call (1 GPU): full GPU suite, the default bench line, ncu --set full captures of the HBM-bound kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/c7_tests.log 2>&1
echo "tests rc=$?" >> $O/c7_tests.log
timeout 600 python bench.py > $O/c7_bench.log 2>&1
echo "bench rc=$?" >> $O/c7_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -f -o $O/r2_kernels \
  -k regex:"conv3x3_halo|wgrad_|bn_bwd_|bn_apply|bn_finalize" python tools/gpu_ncu_kernels.py > $O/c7_ncu.log 2>&1
echo "ncu rc=$?" >> $O/c7_ncu.log
tail -n 22 $O/c7_tests.log
grep -h '^{\|rc=' $O/c7_bench.log | cut -c1-300
tail -n 3 $O/c7_ncu.log
ls -la $O/r2_kernels.ncu-rep
