#!/bin/bash
# Round 2, the last GPU seconds of the round: the primary-bins build with 1..32 spheres per warp, cached rectangles and four atomics in flight.
#  A. tests/test_gpu_lbvh.py (all of it) against the source built with -DRT_PRIMARY_BINS_DEFAULT=1
#  B. off / on timing of configs[2] and configs[3] (shipped library, option set explicitly)
#  C. if time is left: the LBVH tests of tests/test_gpu_shipped_path.py, bins on by default
O=gpurun_out/r02pbins2; mkdir -p $O
T0=$SECONDS
RTB200_LIB=$PWD/_variants/librtb200_pb1.so timeout 42 python -m pytest tests/test_gpu_lbvh.py -m gpu -q --timeout 40 > $O/pytest_lbvh_bins_default_on.log 2>&1; echo "rc=$? t=$((SECONDS-T0))" >> $O/pytest_lbvh_bins_default_on.log; tail -4 $O/pytest_lbvh_bins_default_on.log
timeout 12 python profiles/pbins_timing.py > $O/timing.jsonl 2> $O/timing.err; echo "timing rc=$? t=$((SECONDS-T0))"; cat $O/timing.jsonl
[ $((SECONDS-T0)) -lt 48 ] && RTB200_LIB=$PWD/_variants/librtb200_pb1.so timeout 18 python -m pytest tests/test_gpu_shipped_path.py -m gpu -q --timeout 17 -k "world4_partition or non_finite or two_streams" > $O/pytest_shipped_bins_default_on.log 2>&1; echo "rc=$? t=$((SECONDS-T0))"; tail -3 $O/pytest_shipped_bins_default_on.log
