#!/bin/bash
# Short multi-GPU check (one box, N GPUs): multi-GPU tests, bench.py at N = 2 .. MAXG under torchrun, per-config timings at MAXG.
# usage: bash profiles/scale_quick.sh <max_gpus>
MAXG=${1:-4}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -2 gpurun_out/pytest_multi.log
: > gpurun_out/scale_quick.jsonl
for n in 2 4 8; do
  if [ $n -le $MAXG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 3 >> gpurun_out/scale_quick.jsonl 2>> gpurun_out/scale_quick.err
  fi
done
python - <<'PY'
import json
for l in open("gpurun_out/scale_quick.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print("N=%s" % d.get("n_gpus"), "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d.get("gather_compare"))
PY
python profiles/run_configs.py gpurun_out/configs_final_n$MAXG.json --skip-brute3 --devices=$MAXG 2>>gpurun_out/scale_quick.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('N=%d' % d['n_devices'], d['config'], 'ms', round(d['kernel_ms_best'],3), 'checksum', d['checksum'])
"
