#!/bin/bash
# fused up-conv: unit test vs oracle, whole GPU suite, A/B bench, launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fastpath_kernels.py -q -x --timeout 120 -p no:cacheprovider -k "modconv_up_fused" > gpurun_out/r2a_pytest_upfused.log 2>&1
echo "upfused test exit $?"; tail -15 gpurun_out/r2a_pytest_upfused.log | cut -c1-400
timeout 600 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2a_pytest_gpu.log | cut -c1-300
for f in 1 0; do
  RW_UP_FUSED=$f timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_fused$f.json 2> gpurun_out/r2a_bench_fused$f.err
  echo "bench fused=$f exit $?"; tail -c 300 gpurun_out/r2a_bench_fused$f.err; head -c 400 gpurun_out/r2a_bench_fused$f.json; echo
done
RW_UP_FUSED_MINW=32 timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_minw32.json 2> gpurun_out/r2a_bench_minw32.err
echo "bench minw32 exit $?"; head -c 300 gpurun_out/r2a_bench_minw32.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2a_ncu_bench.log 2>&1
echo "ncu exit $?"; python tools/launch_summary.py gpurun_out/r2a_launches.csv 2>&1 | head -20
