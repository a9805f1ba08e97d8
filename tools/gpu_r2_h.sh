#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_fastpath_kernels.py -q -x --timeout 120 -p no:cacheprovider -k "cta_pair_large or modconv_up_fused" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_holes.py tests/test_gpu_proggan.py tests/test_gpu_config4.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2h_pytest.log | cut -c1-300 | head -20
timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench exit $?"; tail -c 200 gpurun_out/r2h_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1]); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', 'conv frac', round(d['roofline']['frac'],4))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_tc|upconv' -s 26 -c 13 --csv --log-file gpurun_out/r2h_conv_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2h_ncu.log 2>&1
grep -E "gpu__time" gpurun_out/r2h_conv_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | paste -sd' '
B="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 5 -c 1 -f -o gpurun_out/r2h_upfused_l13 $B > gpurun_out/r2h_ncu_l13.log 2>&1; echo "ncu l13 $?"
bash tools/sanitize.sh 2>&1 | tail -8
