#!/bin/bash
# Round-2 thirteenth GPU call: (1) evaluation with conv + BN + residual + ReLU in one launch: eval tests + cfg5 timing
# A/B; (2) co-resident weight-gradient configuration (B200SEG_WGRAD_CORES=1): op tests + step A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
B200SEG_EVAL_FUSED=1 timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "eval or infer or three_scale or full_size_value" > $O/c13_eval_tests.log 2>&1
echo "eval tests rc=$?" >> $O/c13_eval_tests.log
tail -n 6 $O/c13_eval_tests.log | cut -c1-300
for f in 0 1; do B200SEG_EVAL_FUSED=$f timeout 300 python tools/gpu_eval_bench.py 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300; done | tee $O/c13_eval_bench.log
B200SEG_WGRAD_CORES=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py -m gpu -x -q -p no:cacheprovider -k "conv or block or module" > $O/c13_wgrad_tests.log 2>&1
echo "wgrad tests rc=$?" >> $O/c13_wgrad_tests.log
tail -n 6 $O/c13_wgrad_tests.log | cut -c1-300
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-recipe"
export B200SEG_TIME_ONLY=1
for cfg in "B200SEG_WGRAD_CORES=0" "B200SEG_WGRAD_CORES=1" "B200SEG_WGRAD_CORES=0" "B200SEG_WGRAD_CORES=1"; do
  echo "== $cfg"
  env $cfg timeout 200 $B 2>&1 | grep -h '^{\|Error\|error' | cut -c1-300
done | tee $O/c13_wgrad_ab.log
