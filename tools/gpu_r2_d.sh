#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/debug_cov2.py 2>&1 | grep -E "batch|per-image" | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2d_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/r2d_pytest_gpu.log | cut -c1-250 | head -30
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/r2d_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d.get('roofline_cov'))
    for k,v in d['extra'].items(): print(k, json.dumps(v)[:700])
    print('cpu', json.dumps(d.get('cpu_baseline'))[:500])
except Exception as e: print('parse error', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:conv_tc|upconv_fused' -s 39 -c 26 --csv --log-file gpurun_out/r2d_conv_dram.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2d_ncu.log 2>&1
echo "ncu exit $?"; python tools/dram_summary.py gpurun_out/r2d_conv_dram.csv gpurun_out/r2d_dram_per_launch.json "x" 2>&1 | tail -4
