#!/bin/bash
# Round 2: 4-wide LBVH nodes — parity tests on the default build, then configs[2]/[3] timings for the occupancy variants.
O=gpurun_out/r02wide; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_lbvh.py tests/test_gpu_shipped_path.py -m gpu -q -s --timeout 600 -k "lbvh or bins or config4 or non_finite or update or config3 or queries" > $O/pytest_lbvh.log 2>&1; echo "rc=$?" >> $O/pytest_lbvh.log; tail -4 $O/pytest_lbvh.log
for v in default mb6 mb5 mb4; do
  lib=$PWD/_variants/librtb200_$v.so; [ $v = default ] && lib=$PWD/uu-infogr-raytracer_b200/librtb200.so
  for s in config3 config4; do
    echo "$v $s: $(RTB200_LIB=$lib timeout 120 python profiles/prof_driver.py 1 6 $s 2>&1 | tail -2 | tr '\n' ' ')"
  done
done | tee $O/variants.txt
