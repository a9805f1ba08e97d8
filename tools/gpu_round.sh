#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list.  Everything a later step needs is written
# under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; echo
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -20
