#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_proggan.py tests/test_gpu_parity_holes.py tests/test_gpu_config4.py tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2e_pytest.log | cut -c1-300 | head -40
bash tools/sanitize.sh
