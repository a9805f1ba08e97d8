#!/bin/bash
# ncu --set full captures of the layer-13 blur and the layer-13 conv (one forward = 6 blur and
# 13 conv launches; the 6th blur / 12th conv are layer 13)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:blur_up -s 5 -c 1 -f -o gpurun_out/blur_l13 $B > gpurun_out/ncu_blur_l13.log 2>&1; echo "blur $?"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 11 -c 1 -f -o gpurun_out/conv_l13 $B > gpurun_out/ncu_conv_l13.log 2>&1; echo "conv $?"
ls -la gpurun_out/*.ncu-rep
