#!/bin/bash
# Round 2, 4-GPU box: packed gather with 2-D pixel blocks (RGB24 and RGB24 + grey) on real devices: multi-process test + N = 4 benches.
O=gpurun_out/r02m4d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -m gpu -k "packed or ipc" > $O/pytest_multi_packed.log 2>&1; echo "rc=$?" >> $O/pytest_multi_packed.log; tail -3 $O/pytest_multi_packed.log
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{"; }
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('packed_gather'))"; }
RTB200_GATHER_MODE=2 run 4 > $O/bench_n4_mode2.json; show $O/bench_n4_mode2.json
RTB200_GATHER_MODE=1 run 4 > $O/bench_n4_mode1.json; show $O/bench_n4_mode1.json
run 4 > $O/bench_n4_auto.json; show $O/bench_n4_auto.json
