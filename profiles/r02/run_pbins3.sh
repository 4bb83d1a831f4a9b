#!/bin/bash
# Round 2, the last 40 GPU-seconds: the SHIPPED library (primary bins on by default, new build kernel) on the LBVH tests of the other
# GPU test files that run_pbins2.sh did not reach, then smoke().
O=gpurun_out/r02pbins3; mkdir -p $O
T0=$SECONDS
timeout 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_shipped_path.py -m gpu -q --timeout 18 -k "degenerate or in_library_multi_device_equals_single or sparse_d2h_batch" > $O/pytest_rest_lbvh.log 2>&1; echo "rc=$? t=$((SECONDS-T0))" >> $O/pytest_rest_lbvh.log; tail -3 $O/pytest_rest_lbvh.log
[ $((SECONDS-T0)) -lt 19 ] && timeout 7 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$? t=$((SECONDS-T0))" >> $O/smoke.log; tail -2 $O/smoke.log
