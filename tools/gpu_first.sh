#!/bin/bash
# First-contact GPU run: per-amode kernel tests in separate processes (a bad descriptor can
# kill the context), then forward parity.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for m in 0 1 2; do
  UNETB200_TEST_AMODES=$m timeout 300 python -m pytest tests/test_conv_kernels.py -q -x -m gpu -k "not convt and not stem" > gpurun_out/kern_amode$m.log 2>&1
  echo "amode $m exit $?" >> gpurun_out/summary.txt
done
timeout 300 python -m pytest tests/test_conv_kernels.py -q -m gpu -k "convt or stem" > gpurun_out/kern_other.log 2>&1
echo "convt/stem exit $?" >> gpurun_out/summary.txt
for m in 0 1 2; do
  UNETB200_AMODE=$m UNETB200_TEST_AMODES=$m timeout 600 python -m pytest tests/test_forward_parity.py -q -s -m gpu > gpurun_out/fwd_amode$m.log 2>&1
  echo "forward amode $m exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
tail -5 gpurun_out/kern_amode*.log gpurun_out/kern_other.log gpurun_out/fwd_amode*.log
