#!/bin/bash
O=gpurun_out/r02v3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
for s in config3 config4; do echo "$s: $(timeout 120 python profiles/prof_driver.py 1 6 $s 2>&1 | tail -2 | tr '\n' ' ')"; done | tee $O/lbvh.txt
timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2>/dev/null
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('bench', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
