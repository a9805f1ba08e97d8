#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fastpath_kernels.py tests/test_gpu_parity_holes.py tests/test_gpu_parity.py tests/test_gpu_config4.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2i_pytest.log | cut -c1-250 | head -20
timeout 120 python tools/prof_upconv.py 2>&1 | tail -8
timeout 120 python tools/prof_upconv.py 32 512 256 64 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sample-images 0 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cov', d['extra'].get('key_covariance_samples_per_s'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_cov_launches_b250.csv python tools/prof_r2.py cov 250 3 > gpurun_out/r2i_cov_b250.log 2>&1; python tools/launch_summary.py gpurun_out/r2i_cov_launches_b250.csv | head -8
