#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_kernels.py -x -q -m gpu -k "row or head" > gpurun_out/pytest_row.log 2>&1; echo "unit rc=$?"; tail -n 15 gpurun_out/pytest_row.log
timeout 600 python -m pytest tests/test_forward_parity.py -x -q -m gpu -k "row_kernel" > gpurun_out/pytest_row2.log 2>&1; echo "fwd rc=$?"; tail -n 15 gpurun_out/pytest_row2.log
for r in 0 1 3; do
  UNETB200_ROW64=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_row$r.json > gpurun_out/bench_row$r.json 2> gpurun_out/bench_row$r.err
  echo "row64=$r rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_row$r.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms roof',d['roofline']['frac'])" 2>&1 | tail -n 1)"
  python - <<PY
import json
d=json.load(open('gpurun_out/layers_row$r.json'))
print({l['layer']: l['ms'] for l in d['layers'] if l['layer'] in ('down1.net.0','down1.net.3','conv1.net.0','conv1.net.3','up1')})
PY
done
