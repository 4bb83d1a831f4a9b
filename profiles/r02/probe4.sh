mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_shipped_path.py tests/test_gpu_parity.py -q --timeout 600 -m gpu 2>&1 | tail -3
P="python profiles/e2e_probe.py"
( $P; $P; $P --zero-copy 2; $P --sparse 0; RTB200_BANDS=3 $P; RTB200_BANDS=10 $P; $P --w 1280 --h 720; $P --w 1920 --h 1080; $P --w 1920 --h 1080 --zero-copy 0; $P --w 2560 --h 1440 --zero-copy 2; $P --w 2560 --h 1440 --zero-copy 0 ) > gpurun_out/r02/e2e_probe4.jsonl 2>gpurun_out/r02/e2e_probe4.err
tail -3 gpurun_out/r02/e2e_probe4.err
