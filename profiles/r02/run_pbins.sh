#!/bin/bash
# Round 2, last GPU call of the round (3.3 GPU-minutes left): RT_OPT_PRIMARY_BINS on the device.
#  1. the four primary-bins GPU tests (shipped library, option set explicitly)
#  2. off / on timing of BASELINE configs[2] and configs[3] at 4K (shipped library)
#  3. the whole GPU suite against the same source built with -DRT_PRIMARY_BINS_DEFAULT=1 (every LBVH frame of every test uses the bins)
#  4. the PB_CAP = 64 variant's timing, if time is left
O=gpurun_out/r02pbins; mkdir -p $O
T0=$SECONDS
timeout 60 python -m pytest tests/test_gpu_lbvh.py -m gpu -q -k primary_bins --timeout 55 -s > $O/pytest_pbins.log 2>&1; echo "rc=$? t=$((SECONDS-T0))" >> $O/pytest_pbins.log; tail -4 $O/pytest_pbins.log
timeout 25 python profiles/pbins_timing.py > $O/timing_default.jsonl 2> $O/timing_default.err; echo "timing rc=$? t=$((SECONDS-T0))"; cat $O/timing_default.jsonl
RTB200_LIB=$PWD/_variants/librtb200_pb1.so timeout 90 python -m pytest tests -m gpu -q -k "not primary_bins" --timeout 85 > $O/pytest_gpu_bins_default_on.log 2>&1; echo "rc=$? t=$((SECONDS-T0))" >> $O/pytest_gpu_bins_default_on.log; tail -12 $O/pytest_gpu_bins_default_on.log
[ $((SECONDS-T0)) -lt 150 ] && RTB200_LIB=$PWD/_variants/librtb200_cap64.so timeout 20 python profiles/pbins_timing.py > $O/timing_cap64.jsonl 2> $O/timing_cap64.err; echo "cap64 rc=$? t=$((SECONDS-T0))"; cat $O/timing_cap64.jsonl
