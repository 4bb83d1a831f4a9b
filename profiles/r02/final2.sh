#!/bin/bash
# Round 2 evidence run on ONE GPU (final build of round 2): GPU parity suite, smoke, both bench arms, ncu launch list of the bench
# command, one ncu --set full capture each of k_render_tiny (bench workload) and k_render_lbvh (configs[3]).
O=gpurun_out/r02final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
timeout 120 python profiles/prof_driver.py 16 3 > $O/prof_drv.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_tiny -s 2 -c 1 -o $O/prof_tiny -f \
    python profiles/prof_driver.py 16 3 > $O/ncu_prof_tiny.log 2>&1
timeout 120 python profiles/prof_driver.py 1 3 config4 > $O/prof_drv4.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_lbvh -s 2 -c 1 -o $O/prof_lbvh -f \
    python profiles/prof_driver.py 1 3 config4 > $O/ncu_prof_lbvh.log 2>&1
tail -c 400 $O/bench.json; ls -la $O
timeout 600 python profiles/run_configs.py $O/configs_n1.json --skip-brute3 > $O/configs_n1.log 2>&1
