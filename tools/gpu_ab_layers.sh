#!/bin/bash
# per-layer A/B: each arg = "ENV=val,ENV=val" ; prints the whole layer table compactly
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $(echo "$cfg" | tr ',' ' ') timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_ab$i.json > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  echo "[$cfg] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_ab$i.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms')" 2>&1 | tail -n 1)"
  python - <<PY
import json
d=json.load(open('gpurun_out/layers_ab$i.json'))
print(' '.join(f"{l['layer'].replace('.net.','.').replace('bottleneck','bt')}={l['ms']:.3f}" for l in d['layers']), 'sum', round(d['ms_per_step_profiled'],3))
PY
done
