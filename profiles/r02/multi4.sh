#!/bin/bash
# Round 2, 4-GPU box: multi-GPU parity tests on real devices after the last code change, then N = 2 and N = 4 benches per gather mode.
O=gpurun_out/r02m4; mkdir -p $O
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 -m gpu > $O/pytest_multi_n4.log 2>&1; echo "rc=$?" >> $O/pytest_multi_n4.log; tail -5 $O/pytest_multi_n4.log
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | grep "^{"; }
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('packed_gather'), (d.get('gather_compare') or {}).get('nccl_gather_ms_per_step'))"; }
run 2 > $O/bench_n2_auto.json; show $O/bench_n2_auto.json
run 4 > $O/bench_n4_auto.json; show $O/bench_n4_auto.json
RTB200_GATHER_MODE=2 run 4 > $O/bench_n4_mode2.json; show $O/bench_n4_mode2.json
RTB200_GATHER_MODE=0 run 4 > $O/bench_n4_plain.json; show $O/bench_n4_plain.json
RTB200_GATHER_MODE=1 RTB200_SINK_TILES=1 RTB200_PEER_TILES=1 run 4 > $O/bench_n4_mode1_1_1.json; show $O/bench_n4_mode1_1_1.json
RTB200_GATHER_MODE=1 RTB200_SINK_TILES=2 RTB200_PEER_TILES=3 run 4 > $O/bench_n4_mode1_2_3.json; show $O/bench_n4_mode1_2_3.json
