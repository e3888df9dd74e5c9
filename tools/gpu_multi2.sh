#!/bin/bash
# multi-GPU check on one box: launcher tests on all visible GPUs, then the torchrun bench at the given Ns
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_inference_gpu.py -x -q -m gpu -k "launcher or multi_gpu or packed" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_multi.log
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "N=$n rc=$?"; tail -n 3 gpurun_out/bench_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n$n.json'))
    print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'ratio', round(d['e2e']['value']/d['value'],4))
    for k in ('sharded_512','shard_bitident','launcher_threads'): print(k, json.dumps(d.get(k))[:420])
except Exception as e: print('parse error', e)
PY
done
