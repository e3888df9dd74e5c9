#!/bin/bash
mkdir -p gpurun_out
# pair kernels first, in their own process (a protocol bug traps the context)
timeout 600 python -m pytest tests/test_conv_kernels.py -x -q -m gpu -k "test_convt2x2 or test_conv3x3" > gpurun_out/pytest_pair_kern.log 2>&1
echo "kernels rc=$?"; tail -n 15 gpurun_out/pytest_pair_kern.log | cut -c1-300
timeout 600 python -m pytest tests/test_forward_parity.py -x -q -m gpu > gpurun_out/pytest_pair_fwd.log 2>&1
echo "forward rc=$?"; tail -n 5 gpurun_out/pytest_pair_fwd.log | cut -c1-300
i=0
for cfg in "UNETB200_PAIR=2" "UNETB200_PAIR=1" "UNETB200_PAIR=0"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_$i.json > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
  echo "[$i] $cfg rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_$i.json')); print(round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms roof',d['roofline']['frac'],'e2e',round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])" 2>&1 | tail -n 1)"
done
