#!/bin/bash
# round-2 check: GPU tests, smoke, default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
for k in ('value','ms_per_step'): print(k, d[k])
print('e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
for k in ('sharded_512','shard_bitident','launcher_threads','hires_1024','run_unet_batch_1080p','cpu_baseline'): print(k, json.dumps(d.get(k))[:700])
print('lat', json.dumps(d.get('latency_b1'))[:900])
PY
