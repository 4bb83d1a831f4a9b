#!/bin/bash
O=gpurun_out/r02v2; mkdir -p $O
for v in default packxy default packxy; do
  lib=$PWD/_variants/librtb200_$v.so; [ $v = default ] && lib=$PWD/uu-infogr-raytracer_b200/librtb200.so
  RTB200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_$v.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]);print('$v', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done | tee $O/variants.txt
RTB200_LIB=$PWD/_variants/librtb200_packxy.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shipped_path.py -m gpu -q -x --timeout 600 -k "not lbvh and not config4" 2>&1 | tail -2 | tee -a $O/variants.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_tiny -s 2 -c 1 -o $O/prof_tiny -f python profiles/prof_driver.py 16 3 > $O/ncu_prof_tiny.log 2>&1
