#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fastpath_kernels.py -q --timeout 200 -p no:cacheprovider -k "layer_level" 2>&1 | tail -12
