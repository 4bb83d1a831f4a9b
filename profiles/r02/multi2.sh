mkdir -p gpurun_out/r02
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 -m gpu -x > gpurun_out/r02/pytest_multi_n2.log 2>&1; tail -8 gpurun_out/r02/pytest_multi_n2.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 "$@"; }
run > gpurun_out/r02/bench_n2_auto.json 2> gpurun_out/r02/bench_n2_auto.err; tail -2 gpurun_out/r02/bench_n2_auto.err
RTB200_GATHER_MODE=2 run > gpurun_out/r02/bench_n2_packed.json 2> gpurun_out/r02/bench_n2_packed.err; tail -2 gpurun_out/r02/bench_n2_packed.err
RTB200_GATHER_MODE=1 run > gpurun_out/r02/bench_n2_rgb.json 2> gpurun_out/r02/bench_n2_rgb.err
RTB200_GATHER_MODE=2 RTB200_SINK_TILES=9 RTB200_PEER_TILES=10 run > gpurun_out/r02/bench_n2_packed_9_10.json 2> /dev/null
