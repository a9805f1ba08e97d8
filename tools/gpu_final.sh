#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_final.log | cut -c1-300
timeout 120 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; tail -c 300 gpurun_out/bench_final.err; head -c 700 gpurun_out/bench_final.json; echo
