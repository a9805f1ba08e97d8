#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/debug_insert.py 2>&1 | grep -v Warn | tail -5
timeout 600 python -m pytest tests/test_gpu_config4.py tests/test_gpu_fastpath_kernels.py -q --timeout 200 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/r2c_pytest.log | cut -c1-250 | head -30
for f in 1 0; do
  RW_UP_FUSED=$f timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2c_bench_fused$f.json 2> gpurun_out/r2c_bench_fused$f.err
  echo "bench fused=$f exit $?"; tail -c 300 gpurun_out/r2c_bench_fused$f.err; head -c 250 gpurun_out/r2c_bench_fused$f.json; echo
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:upconv -s 6 -c 6 --csv --log-file gpurun_out/r2c_up_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2c_ncu.log 2>&1
grep -E "gpu__time" gpurun_out/r2c_up_launches.csv | awk -F'","' '{print $NF}' | tr -d '"'
