#!/bin/bash
# Round 2, GPU run 1 (1 GPU): full GPU parity suite on the default build, bench, then the same with the RT_GATES_V2 build.
set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02/gpu.txt
nproc >> gpurun_out/r02/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02/pytest_gpu_run1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu_run1.log
tail -5 gpurun_out/r02/pytest_gpu_run1.log
timeout 600 python bench.py > gpurun_out/r02/bench_run1.json 2> gpurun_out/r02/bench_run1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02/bench_run1.err
if [ -f _variants/librtb200_v2.so ]; then
  RTB200_LIB=$PWD/_variants/librtb200_v2.so timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "not multi" > gpurun_out/r02/pytest_gpu_v2.log 2>&1; echo "pytest v2 rc=$?" >> gpurun_out/r02/pytest_gpu_v2.log
  tail -3 gpurun_out/r02/pytest_gpu_v2.log
  RTB200_LIB=$PWD/_variants/librtb200_v2.so timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_v2.json 2> gpurun_out/r02/bench_v2.err; echo "bench v2 rc=$?"
fi
