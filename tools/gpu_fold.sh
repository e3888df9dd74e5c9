#!/bin/bash
# A/B of the folded up-conv levels (option fold_up, bit k = decoder level k): parity tests with every level folded,
# then the per-layer bench for each "ENV=.." config given as an argument
mkdir -p gpurun_out
UNETB200_FOLD_UP=${FOLD_TEST:-15} timeout 900 python -m pytest tests/test_forward_parity.py -x -q -m gpu > gpurun_out/pytest_fold.log 2>&1; echo "pytest(fold) rc=$?"; tail -n 8 gpurun_out/pytest_fold.log
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $(echo "$cfg" | tr ',' ' ') timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_f$i.json > gpurun_out/bench_f$i.json 2> gpurun_out/bench_f$i.err
  echo "[$cfg] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_f$i.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks'])" 2>&1 | tail -n 1)"
  python - <<PY
import json
d=json.load(open('gpurun_out/layers_f$i.json'))
print({l['layer']: l['ms'] for l in d['layers'] if l['layer'].startswith('up') or l['layer'].endswith('net.0') and l['layer'].startswith('conv')}, 'sum', round(d['ms_per_step_profiled'],3))
PY
done
