#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upconv_fused -s 2 -c 1 -o gpurun_out/r2k_upfused_v2 -f python tools/prof_upconv.py > gpurun_out/r2k_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2k_ncu.log
