#!/bin/bash
# tests + short benches; args: "ENV=val,ENV=val[;bench args]" configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs=$(echo "$cfg" | cut -d';' -f1 | tr ',' ' ')
  extra=$(echo "$cfg" | cut -s -d';' -f2)
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $extra \
     --layers-out gpurun_out/layers_$i.json > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
  echo "[$i] $cfg rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_$i.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms roof',d['roofline']['frac'],'e2e',round(d['e2e']['value'],1), 'lat', d.get('latency_b1'), d['clocks']['sm_mhz'], d['clocks']['reasons'])" 2>&1 | tail -n 1)"
done
