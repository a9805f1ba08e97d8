#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fastpath_kernels.py -q --timeout 200 -p no:cacheprovider -k "up_fused or upconv" > gpurun_out/r2j_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2j_pytest.log | cut -c1-250 | head -20
timeout 120 python tools/prof_upconv.py 2>&1 | tail -19
timeout 120 python tools/prof_upconv.py 32 512 256 64 2>&1 | tail -19 | head -3
