mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_shipped_path.py tests/test_gpu_parity.py tests/test_gpu_lbvh.py -q --timeout 600 -m gpu 2>&1 | tail -5
P="python profiles/e2e_probe.py"
( $P --zero-copy 0; $P; $P --sparse 0; $P --precleared 1; RTB200_FILL_THREADS=12 $P; RTB200_FILL_THREADS=3 $P; $P --w 1280 --h 720; $P --w 1280 --h 720 --zero-copy 0; $P --w 7680 --h 4320 --frames 16; $P --w 7680 --h 4320 --frames 16 --zero-copy 0 ) > gpurun_out/r02/e2e_probe3.jsonl 2>gpurun_out/r02/e2e_probe3.err
tail -3 gpurun_out/r02/e2e_probe3.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/r02/bench_zc.json 2>gpurun_out/r02/bench_zc.err; tail -2 gpurun_out/r02/bench_zc.err
