#!/bin/bash
# Round 2: full GPU suite on ONE GPU with the multi-GPU tests mapped onto it (repeated device ids / all ranks on device 0), final library.
O=gpurun_out/r02final3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rs > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench.json 2>/dev/null
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('bench', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
