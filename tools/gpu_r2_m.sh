#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/r2m_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2m_pytest.log | cut -c1-250 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sample-images 0 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cov', d['extra'].get('key_covariance_samples_per_s'), d.get('clocks'))
PY
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2m_launches_all.csv $B > gpurun_out/r2m_ncu_all.log 2>&1; echo "ncu all $?"
python tools/launch_summary.py gpurun_out/r2m_launches_all.csv | head -14
