#!/bin/bash
# Round 2: occupancy re-check AFTER the 2-D pixel blocks: resident CTAs per SM of the tiny and the LBVH kernels (9 / 10 / 12 vs the shipped 8 / 10).
O=gpurun_out/r02occ; mkdir -p $O
for v in default t9 t10 t12; do
  lib=$PWD/_variants/librtb200_$v.so; [ $v = default ] && lib=$PWD/uu-infogr-raytracer_b200/librtb200.so
  RTB200_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-extras --steps 60 > $O/bench_$v.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]);print('$v tiny', round(d['value']), round(d['ms_per_step'],4))"
  for s in config3 config4; do echo "$v $s: $(RTB200_LIB=$lib timeout 100 python profiles/prof_driver.py 1 5 $s 2>&1 | tail -1)"; done
done | tee $O/variants.txt
