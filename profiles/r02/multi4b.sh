#!/bin/bash
# Round 2, 4-GPU box: plain gather at N = 4 and N = 3 with / without the sparse (black spans not sent) option.
O=gpurun_out/r02m4b; mkdir -p $O
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | grep "^{"; }
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('packed_gather'))"; }
run 4 > $O/bench_n4_plain.json; show $O/bench_n4_plain.json
RTB200_SPARSE_MIN_WORLD=2 run 4 > $O/bench_n4_sparse.json; show $O/bench_n4_sparse.json
RTB200_SPARSE_MIN_WORLD=2 run 2 > $O/bench_n2_sparse.json; show $O/bench_n2_sparse.json
run 2 > $O/bench_n2_plain.json; show $O/bench_n2_plain.json
