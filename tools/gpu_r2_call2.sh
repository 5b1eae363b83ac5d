#!/bin/bash
# Round-2 second GPU call (1 GPU): new tests, occupancy report, stream report, bench after the publish restructure.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/gpu_occupancy.py > $O/c2_occupancy.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/c2_tests.log 2>&1
echo "tests rc=$?" >> $O/c2_tests.log
timeout 300 python tools/gpu_stream_report.py > $O/c2_streams.log 2>&1
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline > $O/c2_bench.log 2>&1
tail -n 30 $O/c2_tests.log
grep -h '^{' $O/c2_bench.log | cut -c1-300
tail -n 12 $O/c2_occupancy.log
