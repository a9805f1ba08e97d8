#!/bin/bash
# ncu --set full captures of the layer-13 blur (both variants) and the layer-13 conv
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:blur_up -s 5 -c 1 -f -o gpurun_out/blur_pipe $B > gpurun_out/ncu_blur_pipe.log 2>&1; echo "pipe $?"
RW_BLUR_PIPE=0 timeout 200 ncu --set full --import-source on --clock-control none -k regex:blur_up -s 5 -c 1 -f -o gpurun_out/blur_fused $B > gpurun_out/ncu_blur_fused.log 2>&1; echo "fused $?"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 11 -c 1 -f -o gpurun_out/conv_l13 $B > gpurun_out/ncu_conv_l13.log 2>&1; echo "conv $?"
ls -la gpurun_out/*.ncu-rep
