#!/bin/bash
# First GPU call of a round (one box, ~10 GPU-minutes): confirms main, then the measurements DESIGN.md still lacks.
#   gpurun --timeout 1200 -- 'bash tools/gpu_first_call.sh'
# Everything lands in gpurun_out/first_*.log; nothing here is a bench value to report except first_bench.log's JSON line.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/first_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/first_tests.log
python bench.py > gpurun_out/first_bench.log 2>&1
python bench.py --no-cpu-baseline --torch-gpu-baseline --steps 10 > gpurun_out/first_bench_torch.log 2>&1
python tools/gpu_stream_report.py > gpurun_out/first_streams.log 2>&1
# launch list (two eager steps: warm-up + one inside the "timed_step" NVTX range; shares of the kernels and the input of
# tools/conv_classes.py, which takes the first step) - never a timing source for bench values
B200SEG_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/first_launches.csv python bench.py --no-graph --no-cpu-baseline > gpurun_out/first_ncu.log 2>&1
gzip -f gpurun_out/first_launches.csv
# the same with every ABI call logged: exact attribution of kernel time to convolution classes
# (python -O tools/conv_classes.py gpurun_out/abi_launches.csv.gz --abi gpurun_out/abi_calls.json)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/abi_launches.csv \
  python tools/gpu_abi_log.py > gpurun_out/first_abi.log 2>&1
gzip -f gpurun_out/abi_launches.csv
tail -n 3 gpurun_out/first_tests.log
grep -h '^{' gpurun_out/first_bench.log gpurun_out/first_bench_torch.log | cut -c1-600
