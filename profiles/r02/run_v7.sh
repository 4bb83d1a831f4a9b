#!/bin/bash
O=gpurun_out/r02v7; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
for s in config3 config4; do echo "tile2d $s: $(timeout 120 python profiles/prof_driver.py 1 6 $s 2>&1 | tail -2 | tr '\n' ' ')"; done | tee $O/lbvh.txt
for s in config3 config4; do echo "linear $s: $(RTB200_NO_TILE2D=1 timeout 120 python profiles/prof_driver.py 1 6 $s 2>&1 | tail -2 | tr '\n' ' ')"; done | tee -a $O/lbvh.txt
