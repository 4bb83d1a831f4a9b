mkdir -p gpurun_out/r02
P="python profiles/e2e_probe.py"
( $P --sparse 0; $P; $P --precleared 1; RTB200_RECT_MAX_FRAC=0 $P; RTB200_RECT_MAX_FRAC=50 $P; RTB200_BANDS=2 $P; RTB200_BANDS=4 $P; RTB200_BANDS=12 $P; RTB200_BANDS=16 $P; RTB200_BANDS=1 $P; RTB200_BANDS=1 $P --sparse 0 ) > gpurun_out/r02/e2e_probe1.jsonl 2>gpurun_out/r02/e2e_probe1.err
tail -3 gpurun_out/r02/e2e_probe1.err
