#!/bin/bash
# Round 2: packed two-light Phong pass — full GPU parity suite, then the bench line.
O=gpurun_out/r02pair; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02pair/bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'single', round(d['value_single_frame']['value']), 'gates_off', round(d['value_gates_off']['value']), 'moving', round(d['value_moving_camera']['value']))
for k,v in d['per_config'].items(): print(k, v.get('kernel_ms'), v.get('frame_equals_instrumented', v.get('checksum')))
PY
