#!/bin/bash
# Round-2 eighth GPU call (2 GPUs): SyncBN on the full branch-stream program - captured parity test + full-size benches.
set -u
mkdir -p gpurun_out
O=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "syncbn and True" --durations=3 > $O/c8_tests.log 2>&1
echo "tests rc=$?" >> $O/c8_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 280 $T --master-port 29551 bench.py --gpus 2 --steps 12 --warmup 4 --syncbn --no-cpu-baseline > $O/c8_n2_syncbn.log 2>&1
echo "syncbn rc=$?" >> $O/c8_n2_syncbn.log
B200SEG_SYNCBN_BRANCH_STREAMS=0 timeout 280 $T --master-port 29552 bench.py --gpus 2 --steps 12 --warmup 4 --syncbn --no-cpu-baseline > $O/c8_n2_syncbn_2chain.log 2>&1
echo "syncbn 2chain rc=$?" >> $O/c8_n2_syncbn_2chain.log
tail -n 8 $O/c8_tests.log
for f in $O/c8_n2_syncbn.log $O/c8_n2_syncbn_2chain.log; do grep -h '^{\|rc=\|^rank [0-9]:' $f | cut -c1-220; done
