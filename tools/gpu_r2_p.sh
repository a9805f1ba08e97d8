#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2p_pytest.log | cut -c1-300 | head -30
timeout 300 python tools/bench_modconv.py > gpurun_out/r2p_modconv.json 2> gpurun_out/r2p_modconv.err; echo "modconv exit $?"; tail -c 900 gpurun_out/r2p_modconv.json; tail -3 gpurun_out/r2p_modconv.err
