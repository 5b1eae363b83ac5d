#!/bin/bash
# Round-2 fourth GPU call (1 GPU): validates the staged halo epilogue, the in-launch BN finalisation and the publish path;
# A/B benches.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c4_tests.log 2>&1
echo "tests rc=$?" >> $O/c4_tests.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline > $O/c4_bench.log 2>&1
B200SEG_HALO_FAST=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline > $O/c4_bench_nofast.log 2>&1
B200SEG_FUSED_BN=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline > $O/c4_bench_nofused.log 2>&1
timeout 300 python tools/gpu_stream_report.py > $O/c4_streams.log 2>&1
tail -n 25 $O/c4_tests.log
for f in $O/c4_bench.log $O/c4_bench_nofast.log $O/c4_bench_nofused.log; do grep -h '^{' $f | cut -c1-200; tail -n 3 $f | cut -c1-300; done
grep -A8 "^class " $O/c4_streams.log
