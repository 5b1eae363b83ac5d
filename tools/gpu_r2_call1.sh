#!/bin/bash
# Round-2 first GPU call: parity figures, GPU tests, co-residency A/B, torch-GPU baseline, launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.log 2>&1
timeout 300 python tools/gpu_occupancy.py > $O/c1_occupancy.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/c1_tests.log 2>&1
echo "tests rc=$?" >> $O/c1_tests.log
timeout 600 python tests/diag/gpu_parity_report.py $O/c1_parity.json > $O/c1_parity.log 2>&1
timeout 600 python bench.py --steps 20 > $O/c1_bench_on.log 2>&1
B200SEG_CORESIDENT=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline > $O/c1_bench_off.log 2>&1
B200SEG_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/c1_launches.csv python bench.py --no-graph --no-cpu-baseline --no-torch-gpu-baseline > $O/c1_ncu.log 2>&1
gzip -f $O/c1_launches.csv
tail -n 5 $O/c1_tests.log
grep -h '^{' $O/c1_bench_on.log $O/c1_bench_off.log | cut -c1-400
tail -n 3 $O/c1_occupancy.log
