#!/bin/bash
O=gpurun_out/r02v9; mkdir -p $O
for v in default rows8 default rows8; do
  lib=$PWD/_variants/librtb200_$v.so; [ $v = default ] && lib=$PWD/uu-infogr-raytracer_b200/librtb200.so
  RTB200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_$v.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]);print('$v', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done | tee $O/variants.txt
