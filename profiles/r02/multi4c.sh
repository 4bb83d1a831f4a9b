#!/bin/bash
# Round 2, 4-GPU box: every BASELINE.json config through the in-library multi-device context on 2 and 4 GPUs (final build), e2e multi-device.
O=gpurun_out/r02m4c; mkdir -p $O
timeout 400 python profiles/run_configs.py $O/configs_n2.json --skip-brute3 --devices=2 > $O/configs_n2.log 2>&1
timeout 400 python profiles/run_configs.py $O/configs_n4.json --skip-brute3 --devices=4 > $O/configs_n4.log 2>&1
python - <<'PY'
import json
for n in (2,4):
    for c in json.load(open('gpurun_out/r02m4c/configs_n%d.json' % n)): print(n, c['config'], round(c['kernel_ms_best'],4), c['checksum'])
PY
