#!/bin/bash
# Round-2 third GPU call (2 GPUs): full GPU suite incl. the multi-GPU tests, SyncBN bench at N=2, TMEM occupancy probe.
set -u
mkdir -p gpurun_out
O=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
tools/bin/occ_probe > $O/c3_probe.log 2>&1
B200SEG_TEST_SYNCBN=1 CUDA_MODULE_LOADING=EAGER timeout 900 python -m pytest tests -m gpu -q > $O/c3_tests.log 2>&1
echo "tests rc=$?" >> $O/c3_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 15 --warmup 4 --syncbn --no-cpu-baseline > $O/c3_bench_n2_syncbn.log 2>&1
echo "syncbn bench rc=$?" >> $O/c3_bench_n2_syncbn.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 2 --steps 15 --warmup 4 --no-cpu-baseline > $O/c3_bench_n2.log 2>&1
echo "bench rc=$?" >> $O/c3_bench_n2.log
cat $O/c3_probe.log
tail -n 12 $O/c3_tests.log
grep -h '^{\|rc=' $O/c3_bench_n2_syncbn.log $O/c3_bench_n2.log | cut -c1-260
