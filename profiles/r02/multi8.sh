N=$1; shift
mkdir -p gpurun_out/r02
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | grep "^{"; }
RTB200_GATHER_MODE=0 run > gpurun_out/r02/bench_n${N}_plain.json
run > gpurun_out/r02/bench_n${N}_auto.json
RTB200_SINK_TILES=1 RTB200_PEER_TILES=1 run > gpurun_out/r02/bench_n${N}_packed_1_1.json
RTB200_SINK_TILES=0 RTB200_PEER_TILES=1 run > gpurun_out/r02/bench_n${N}_packed_0_1.json
RTB200_SINK_TILES=1 RTB200_PEER_TILES=3 run > gpurun_out/r02/bench_n${N}_packed_1_3.json
RTB200_GATHER_MODE=1 run > gpurun_out/r02/bench_n${N}_rgb_auto.json
