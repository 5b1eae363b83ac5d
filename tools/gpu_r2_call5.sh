#!/bin/bash
# Round-2 fifth GPU call (1 GPU): train.py boundary test, op tests after the fence change, fused-BN A/B (two rounds).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train_py.py tests/test_gpu_ops.py -m gpu -q > $O/c5_tests.log 2>&1
echo "tests rc=$?" >> $O/c5_tests.log
B="python bench.py --steps 20 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
for r in 1 2; do
  B200SEG_FUSED_BN=1 timeout 200 $B > $O/c5_fused_$r.log 2>&1
  B200SEG_FUSED_BN=0 timeout 200 $B > $O/c5_unfused_$r.log 2>&1
done
tail -n 15 $O/c5_tests.log
for f in $O/c5_fused_1.log $O/c5_unfused_1.log $O/c5_fused_2.log $O/c5_unfused_2.log; do
python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'device %.2f e2e %.2f blocking %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['blocking_read']['ms_per_step']))
PY
done
