#!/bin/bash
# Round 2: device-built shadow bins with the per-sphere along-light pad — parity tests, per-config timings, cells-per-sphere sweep.
O=gpurun_out/r02bins; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_lbvh.py tests/test_gpu_shipped_path.py -m gpu -q -s --timeout 600 -k "lbvh or bins or config4 or non_finite or update or config3" > $O/pytest_lbvh.log 2>&1; echo "rc=$?" >> $O/pytest_lbvh.log; tail -5 $O/pytest_lbvh.log; grep "rt_update_spheres" $O/pytest_lbvh.log
timeout 600 python profiles/run_configs.py $O/configs.json --skip-brute3 > $O/configs.log 2>&1
python - <<'PY'
import json
for c in json.load(open('gpurun_out/r02bins/configs.json')): print(c['config'], round(c['kernel_ms_best'],3), c['checksum'], round(c['scene_upload_s'],3), c['counters'].get('sphere_tests'))
PY
for cps in 1 2 8 16; do
  for s in config3 config4; do
    echo "cells_per_sphere=$cps $s: $(RTB200_SG_CELLS_PER_SPHERE=$cps timeout 120 python profiles/prof_driver.py 1 5 $s 2>&1 | tail -2 | tr '\n' ' ')"
  done
done | tee $O/cells_sweep.txt
echo "host bins: $(RTB200_SG_HOST=1 timeout 120 python profiles/prof_driver.py 1 4 config4 2>&1 | tail -1)" | tee -a $O/cells_sweep.txt
