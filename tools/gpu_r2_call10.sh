#!/bin/bash
# Round-2 tenth GPU call: DeepLabV3+ gradient diagnostic (product vs oracle vs the oracle's one-ulp noise floor) and the
# launch / dependency floor of the captured two-scale step (B200SEG_DRY: every library launch replaced by an empty kernel).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tests/diag/gpu_deepv3_diag.py > $O/c10_deepv3_diag.log 2>&1
echo "diag rc=$?" >> $O/c10_deepv3_diag.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
export B200SEG_TIME_ONLY=1
timeout 200 $B > $O/c10_real.log 2>&1; echo "rc=$?" >> $O/c10_real.log
B200SEG_DRY=1 timeout 200 $B > $O/c10_dry1.log 2>&1; echo "rc=$?" >> $O/c10_dry1.log
B200SEG_DRY=2 timeout 200 $B > $O/c10_dry2.log 2>&1; echo "rc=$?" >> $O/c10_dry2.log
B200SEG_PDL=0 timeout 200 $B > $O/c10_nopdl.log 2>&1; echo "rc=$?" >> $O/c10_nopdl.log
B200SEG_PDL=0 B200SEG_DRY=1 timeout 200 $B > $O/c10_dry1_nopdl.log 2>&1; echo "rc=$?" >> $O/c10_dry1_nopdl.log
head -c 6000 $O/c10_deepv3_diag.log
for f in real dry1 dry2 nopdl dry1_nopdl; do echo "== $f"; grep -h '^{\|rc=\|Error\|error' $O/c10_$f.log | cut -c1-300 | tail -4; done
