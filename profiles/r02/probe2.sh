mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_shipped_path.py tests/test_gpu_parity.py -q --timeout 600 -m gpu 2>&1 | tail -3
P="python profiles/e2e_probe.py"
( $P --sparse 0; $P; $P --precleared 1; RTB200_NO_HEAD_BAND=1 $P; RTB200_BANDS=3 $P; RTB200_BANDS=4 $P; RTB200_BANDS=8 $P; RTB200_FILL_THREADS=12 $P;  $P --w 1280 --h 720; $P --w 1280 --h 720 --sparse 0 ) > gpurun_out/r02/e2e_probe2.jsonl 2>gpurun_out/r02/e2e_probe2.err
tail -3 gpurun_out/r02/e2e_probe2.err
