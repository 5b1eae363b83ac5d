#!/bin/bash
# Round-2 sixth GPU call (2 GPUs): SyncBN (post + one-warp waiter) parity test and full-size bench, all-reduce variants.
set -u
mkdir -p gpurun_out
O=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/c6_tests.log 2>&1
echo "tests rc=$?" >> $O/c6_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 bench.py --gpus 2 --steps 12 --warmup 4 --syncbn --no-cpu-baseline > $O/c6_n2_syncbn.log 2>&1
echo "syncbn rc=$?" >> $O/c6_n2_syncbn.log
timeout 240 $T --master-port 29542 bench.py --gpus 2 --steps 12 --warmup 4 --no-syncbn --no-cpu-baseline > $O/c6_n2_plain.log 2>&1
echo "plain rc=$?" >> $O/c6_n2_plain.log
B200SEG_BUCKET_ALLREDUCE=1 timeout 240 $T --master-port 29543 bench.py --gpus 2 --steps 12 --warmup 4 --no-syncbn --no-cpu-baseline > $O/c6_n2_bucket.log 2>&1
echo "bucket rc=$?" >> $O/c6_n2_bucket.log
tail -n 8 $O/c6_tests.log
for f in $O/c6_n2_syncbn.log $O/c6_n2_plain.log $O/c6_n2_bucket.log; do grep -h '^{\|rc=\|^rank [0-9]:' $f | cut -c1-220; done
