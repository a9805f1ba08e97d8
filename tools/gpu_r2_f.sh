#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RW_UP_CPT=4 timeout 120 python -m pytest tests/test_gpu_fastpath_kernels.py -q -x --timeout 100 -p no:cacheprovider -k "modconv_up_fused" 2>&1 | tail -2
if [ $? -ne 0 ]; then echo "cpt4 unit test failed"; fi
timeout 600 python -m pytest tests/test_gpu_parity_holes.py tests/test_gpu_parity.py tests/test_gpu_proggan.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2f_pytest.log | cut -c1-300 | head -20
for c in 4 8; do
  RW_UP_CPT=$c timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_cpt$c.json 2> gpurun_out/r2f_bench_cpt$c.err
  echo "bench cpt=$c exit $?"; tail -c 200 gpurun_out/r2f_bench_cpt$c.err; head -c 220 gpurun_out/r2f_bench_cpt$c.json; echo
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upconv -s 6 -c 6 --csv --log-file gpurun_out/r2f_up_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2f_ncu.log 2>&1
grep -E "gpu__time" gpurun_out/r2f_up_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 5 -c 1 -f -o gpurun_out/r2f_upfused_l13 $B > gpurun_out/r2f_ncu_l13.log 2>&1; echo "ncu l13 $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_cov_launches_b250.csv python tools/prof_r2.py cov 250 3 > gpurun_out/r2f_cov_b250.log 2>&1; echo "cov launches $?"
python tools/launch_summary.py gpurun_out/r2f_cov_launches_b250.csv | head -12
