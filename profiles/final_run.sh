#!/bin/bash
# Single-GPU evidence run: GPU tests, smoke, bench (both arms), ncu launch list of the bench command, one ncu --set full capture of
# the render kernel, per-config timings. Everything lands in gpurun_out/ (copy what is to be kept into profiles/rNN/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python profiles/prof_driver.py 16 3 > gpurun_out/prof_drv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_render_tiny -s 2 -c 1 -o gpurun_out/prof_final -f \
    python profiles/prof_driver.py 16 3 > gpurun_out/ncu_prof.log 2>&1
python profiles/run_configs.py gpurun_out/configs_final.json > gpurun_out/configs.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('bench', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],1), d['clocks'])
r=json.loads(open('gpurun_out/bench_reference_arm.json').read().strip().splitlines()[-1]); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
for c in json.load(open('gpurun_out/configs_final.json')): print(c['config'], round(c['kernel_ms_best'],3), c['checksum'])
"
