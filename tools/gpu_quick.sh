#!/bin/bash
# args: pytest -k expression ; then per-layer bench for each "ENV=.." config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_quick.log
shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $(echo "$cfg" | tr ',' ' ') timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_q$i.json > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err
  echo "[$cfg] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_q$i.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks'])" 2>&1 | tail -n 1)"
  python - <<PY
import json
d=json.load(open('gpurun_out/layers_q$i.json'))
print({l['layer']: l['ms'] for l in d['layers'] if l['layer'] in ('down1.net.0','down1.net.3','conv1.net.0','conv1.net.3','up1','up2')}, 'sum', round(d['ms_per_step_profiled'],3))
PY
done
