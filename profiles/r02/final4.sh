#!/bin/bash
# Round 2, last library of the round (rt_tiles.cuh refactor): full GPU suite on one GPU + the full bench line.
O=gpurun_out/r02final4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('bench', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'single', round(d['value_single_frame']['value']), [ (k[:12], round(v['kernel_ms'],3)) for k,v in d['per_config'].items()])"
