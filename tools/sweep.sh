#!/bin/bash
# amode x bn_max sweep of the batch-64 bench (per-layer tables land in gpurun_out/)
mkdir -p gpurun_out
for bn in 128 256; do
 for m in 2 1 0; do
  UNETB200_AMODE=$m UNETB200_BN_MAX=$bn timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
     --layers-out gpurun_out/layers_a${m}_bn${bn}.json > gpurun_out/bench_a${m}_bn${bn}.json 2> gpurun_out/bench_a${m}_bn${bn}.err
  echo "amode=$m bn=$bn rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_a${m}_bn${bn}.json')); print(round(d['value'],1),'img/s', d['ms_per_step'],'ms', 'roof',d['roofline']['frac'], 'e2e', round(d['e2e']['value'],1), d['clocks'])" 2>&1 | tail -1)"
 done
done
