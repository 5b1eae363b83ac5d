#!/bin/bash
# Round-2 nineteenth GPU call: the bench line after moving the one-time staging allocation out of the end-to-end timed regions.
set -u
mkdir -p gpurun_out
B200SEG_E2E_DEBUG=1 timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe > gpurun_out/c19_bench.log 2>&1
echo "rc=$?" >> gpurun_out/c19_bench.log
grep -h 'e2e-debug\|rc=\|Error' gpurun_out/c19_bench.log | cut -c1-300
grep -h '^{' gpurun_out/c19_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); e = d['e2e']
print(dict(step=round(d['ms_per_step'], 2), e2e=round(e['ms_per_step'], 2), blocking=round(e['blocking_read']['ms_per_step'], 2), value=d['value'], e2e_value=e['value']))"
