#!/bin/bash
# Multi-GPU scaling run (one box, N GPUs): tests, bench.py at N = 1,2,4,8 under torchrun, configs 4/5 with the in-library
# multi-device context. usage: bash profiles/scale_run.sh <max_gpus>
MAXG=${1:-8}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -3 gpurun_out/pytest_multi.log
: > gpurun_out/scale.jsonl
python bench.py --impl reference --steps 3 --warmup 1 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python bench.py --gpus 1 --steps 30 --warmup 3 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
for n in 2 4 8; do
  if [ $n -le $MAXG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 30 --warmup 3 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
  fi
done
python - <<'PY'
import json
for l in open("gpurun_out/scale.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d.get("impl", "b200"), "N=%s" % d.get("n_gpus"), "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d.get("gather_compare"))
PY
for n in 1 2 4 8; do
  if [ $n -le $MAXG ]; then
    python profiles/run_configs.py gpurun_out/configs45_n$n.json --skip-brute3 --devices=$n 2>>gpurun_out/scale.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('N=%d' % d['n_devices'], d['config'], 'ms', round(d['kernel_ms_best'],3), 'checksum', d['checksum'])
"
  fi
done
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log; tail -3 gpurun_out/pytest_gpu_all.log
python profiles/e2e_multi.py $MAXG | tee gpurun_out/e2e_multi.txt
