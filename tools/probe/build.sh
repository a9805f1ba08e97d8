#!/bin/bash
# builds the hardware probe next to its source (binary is git-ignored)
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probe_sm100 probe_sm100.cu
