#!/bin/bash
# Round-2 seventeenth GPU call (2 GPUs): the N=2 bench line as the driver launches it (SyncBN is the default at N > 1) with this
# round's kernels (narrow BN grids, late PDL trigger, co-resident weight-gradient CTAs), then the captured SyncBN parity test.
set -u
mkdir -p gpurun_out
O=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29561 bench.py --gpus 2 --steps 12 --warmup 4 --no-cpu-baseline > $O/c17_n2.log 2>&1
echo "n2 rc=$?" >> $O/c17_n2.log
grep -h '^{\|rc=' $O/c17_n2.log | cut -c1-600
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider -k "syncbn and True" --durations=3 > $O/c17_tests.log 2>&1
echo "tests rc=$?" >> $O/c17_tests.log
tail -n 6 $O/c17_tests.log | cut -c1-300
