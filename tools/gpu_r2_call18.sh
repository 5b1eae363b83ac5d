#!/bin/bash
# Round-2 eighteenth GPU call: why is the pipelined end-to-end loop (DevicePrefetcher) slower than the blocking one in call 16?
# The same bench line under: default, early PDL trigger, separate BN finalisers; then ncu --set full of the branch kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
for cfg in "X=0" "B200SEG_LIB_VARIANT=early" "B200SEG_BN_CELLS=0" "B200SEG_E2E_DEBUG=1"; do
  echo "== $cfg"
  env $cfg timeout 120 $B 2>&1 | grep -h '^{\|Error\|error\|e2e-debug' | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); e = d['e2e']
        print(dict(step=round(d['ms_per_step'], 2), e2e=round(e['ms_per_step'], 2), blocking=round(e['blocking_read']['ms_per_step'], 2)))
    else:
        print(line.rstrip()[:300])
"
done | tee $O/c18_e2e.log
timeout 170 ncu --set full --clock-control none --import-source on -f -o $O/r2_kernels_final \
  -k regex:"conv3x3_halo|wgrad_|bn_bwd_|bn_apply|aug_" python tools/gpu_ncu_kernels.py > $O/c18_ncu_full.log 2>&1
echo "ncu full rc=$?" >> $O/c18_ncu_full.log
ls -la $O/r2_kernels_final.ncu-rep
