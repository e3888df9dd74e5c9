#!/bin/bash
# unit + forward tests of the phase-stacked level-1 kernel, then an in-process A/B of option fold_stack
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_up.py tests/test_forward_parity.py -x -q -m gpu -k "phase_stacked" > gpurun_out/pytest_stack.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_stack.log
timeout 300 python tools/ab.py fold_stack=0,1 --layers > gpurun_out/ab_stack.log 2>&1; echo "ab rc=$?"; tail -n 30 gpurun_out/ab_stack.log
