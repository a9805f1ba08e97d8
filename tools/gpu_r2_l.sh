#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python tools/prof_upconv.py 2>&1 | tail -19
timeout 120 python tools/prof_upconv.py 32 512 256 64 2>&1 | tail -19 | head -3
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/r2l_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2l_pytest.log | cut -c1-250 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sample-images 0 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cov', d['extra'].get('key_covariance_samples_per_s'))
print(d.get('roofline_upconv'))
PY
