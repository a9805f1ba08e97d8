#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 800 python -m pytest tests/test_gpu_overfit.py tests/test_gpu_config4.py tests/test_gpu_parity.py tests/test_gpu_parity_holes.py -q --timeout 500 -p no:cacheprovider > gpurun_out/r2s_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2s_pytest.log | cut -c1-300 | head -30
