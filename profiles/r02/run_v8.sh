#!/bin/bash
O=gpurun_out/r02v8; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
for m in tile2d linear tile2d linear; do
  [ $m = linear ] && export RTB200_NO_TILE2D=1 || unset RTB200_NO_TILE2D
  timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_$m.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/bench_$m.json').read().strip().splitlines()[-1]);print('$m', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done | tee $O/variants.txt
