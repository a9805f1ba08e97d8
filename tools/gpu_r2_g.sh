#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fastpath_kernels.py -q -x --timeout 120 -p no:cacheprovider -k "256_column or modconv_up_fused" 2>&1 | tail -3
RW_UP_CPT=4 timeout 120 python -m pytest tests/test_gpu_fastpath_kernels.py -q -x --timeout 100 -p no:cacheprovider -k "modconv_up_fused" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_holes.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2g_pytest.log | cut -c1-300 | head -20
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_$name.json 2> gpurun_out/r2g_bench_$name.err
  echo "bench $name exit $?"; tail -c 200 gpurun_out/r2g_bench_$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2g_bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', 'conv frac', round(d['roofline']['frac'],4))"
}
run cpt8_bn256 RW_UP_CPT=8 RW_CONV_BN256=1
run cpt4_bn256 RW_UP_CPT=4 RW_CONV_BN256=1
run cpt8_bn128 RW_UP_CPT=8 RW_CONV_BN256=0
RW_UP_CPT=8 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_tc|upconv' -s 26 -c 13 --csv --log-file gpurun_out/r2g_conv_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2g_ncu.log 2>&1
grep -E "gpu__time" gpurun_out/r2g_conv_launches.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | cut -c1-60 | paste -sd' ' | fold -w 200
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
RW_UP_CPT=8 timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 5 -c 1 -f -o gpurun_out/r2g_upfused_l13 $B > gpurun_out/r2g_ncu_l13.log 2>&1; echo "ncu l13 $?"
RW_UP_CPT=8 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 5 -c 1 -f -o gpurun_out/r2g_conv_l12 $B > gpurun_out/r2g_ncu_l12.log 2>&1; echo "ncu l12 $?"
