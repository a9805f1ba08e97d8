#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2b_pytest_gpu.log | cut -c1-300
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 5 -c 1 -f -o gpurun_out/r2b_upfused_l13 $B > gpurun_out/r2b_ncu_l13.log 2>&1; echo "ncu l13 $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 3 -c 1 -f -o gpurun_out/r2b_upfused_l9 $B > gpurun_out/r2b_ncu_l9.log 2>&1; echo "ncu l9 $?"
ls -la gpurun_out/*.ncu-rep | tail -3
