#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_overfit.py -q --timeout 500 -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error|Warning" gpurun_out/r2r_pytest.log | cut -c1-300 | head -30
timeout 500 python tools/debug_overfit.py 2>&1 | grep "apply_overfit"
