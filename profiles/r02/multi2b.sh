mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_shipped_path.py -q --timeout 600 -m gpu -x -k "packed" > gpurun_out/r02/pytest_packed_n2.log 2>&1; tail -4 gpurun_out/r02/pytest_packed_n2.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 "$@" 2>/dev/null | grep "^{"; }
run > gpurun_out/r02/bench_n2_auto.json
RTB200_GATHER_MODE=2 run > gpurun_out/r02/bench_n2_packed.json
RTB200_GATHER_MODE=2 RTB200_SINK_TILES=9 RTB200_PEER_TILES=10 run > gpurun_out/r02/bench_n2_packed_9_10.json
RTB200_GATHER_MODE=2 RTB200_SINK_TILES=4 RTB200_PEER_TILES=5 run > gpurun_out/r02/bench_n2_packed_4_5.json
RTB200_GATHER_MODE=1 RTB200_SINK_TILES=9 RTB200_PEER_TILES=10 run > gpurun_out/r02/bench_n2_rgb_9_10.json
