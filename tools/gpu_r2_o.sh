#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_overfit.py -q --timeout 500 -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2o_pytest.log | cut -c1-300 | head -40
