#!/bin/bash
# Round-2 sixteenth GPU call (1 GPU): final evidence of the round. Stream-priority A/B, the default bench line (as the driver
# runs it), launch list of one step, ncu --set full of the 48-channel branch kernels + the input pipeline kernels, the
# kernel-class stream report, augment tests + timing, smoke().
set -u
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
for cfg in "B200SEG_PRIO=0" "B200SEG_PRIO=1" "B200SEG_PRIO=1"; do
  echo "== $cfg"
  env B200SEG_TIME_ONLY=1 $cfg timeout 200 $B 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300
done | tee $O/c16_prio_ab.log
timeout 900 python bench.py > $O/c16_bench.log 2>&1
echo "bench rc=$?" >> $O/c16_bench.log
grep -h '^{' $O/c16_bench.log | cut -c1-400
timeout 120 python bench.py --impl reference --steps 1 --warmup 1 > $O/c16_bench_ref.log 2>&1
grep -h '^{' $O/c16_bench_ref.log | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_augment.py -m gpu -q -p no:cacheprovider > $O/c16_augment.log 2>&1
echo "augment rc=$?" >> $O/c16_augment.log
tail -n 3 $O/c16_augment.log
timeout 120 python tests/diag/gpu_augment_bench.py 2>&1 | grep -h '^{' | cut -c1-500 | tee $O/c16_augment_bench.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c16_smoke.log 2>&1
tail -n 2 $O/c16_smoke.log
B200SEG_PROFILE=1 timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/c16_launches.csv python bench.py --no-graph --no-cpu-baseline --no-torch-gpu-baseline --no-recipe > $O/c16_ncu_list.log 2>&1
echo "launch list rc=$?" >> $O/c16_ncu_list.log
gzip -f $O/c16_launches.csv
ls -la $O/c16_launches.csv.gz
