#!/bin/bash
# round-2 validation visit: whole GPU suite, smoke, full bench, launch lists, ncu captures
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r2v2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2v2_pytest_gpu_final.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED|Error" gpurun_out/r2v2_pytest_gpu_final.log | cut -c1-250 | head -20
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v2_bench_final_n1.json 2> gpurun_out/r2v2_bench_final_n1.err
echo "bench exit $?"; tail -c 300 gpurun_out/r2v2_bench_final_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2v2_bench_final_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'up', d.get('roofline_upconv',{}).get('frac'), 'cov', d.get('roofline_cov',{}).get('frac'))
    for k,v in d['extra'].items(): print(k, json.dumps(v)[:300])
except Exception as e: print('parse error', e)
PY
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2v2_launches_all.csv $B > gpurun_out/r2v2_ncu_all.log 2>&1; echo "ncu all $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:conv_tc|upconv_fused' -s 39 -c 26 --csv --log-file gpurun_out/r2v2_conv_dram.csv $B > gpurun_out/r2v2_ncu_conv.log 2>&1; echo "ncu conv $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:upconv_fused -s 5 -c 1 -f -o gpurun_out/r2v2_upfused_l13_final $B > gpurun_out/r2v2_ncu_l13.log 2>&1; echo "ncu l13 $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 5 -c 1 -f -o gpurun_out/r2v2_conv_l12_final $B > gpurun_out/r2v2_ncu_l12.log 2>&1; echo "ncu l12 $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gram_tc -s 2 -c 1 -f -o gpurun_out/r2v2_gram_b250 python tools/prof_r2.py cov 250 3 > gpurun_out/r2v2_ncu_gram.log 2>&1; echo "gram $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:reduce_partials -s 2 -c 1 -f -o gpurun_out/r2v2_reduce_b250 python tools/prof_r2.py cov 250 3 > gpurun_out/r2v2_ncu_reduce.log 2>&1; echo "reduce $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v2_cov_launches_b250.csv python tools/prof_r2.py cov 250 3 > gpurun_out/r2v2_cov_b250.log 2>&1; echo "cov launches $?"
ls -la gpurun_out/*final*.ncu-rep gpurun_out/r2v2_gram_b250.ncu-rep 2>/dev/null | tail -4
timeout 120 python tools/prof_upconv.py > gpurun_out/r2v2_upconv_phases_l13.txt 2>&1; tail -18 gpurun_out/r2v2_upconv_phases_l13.txt | head -4
timeout 120 python tools/prof_upconv.py 32 512 256 64 > gpurun_out/r2v2_upconv_phases_l11.txt 2>&1
for i in 1 2 3 4; do timeout 25 tools/probe/probe_sm100 $i; done > gpurun_out/r2v2_probe_sm100.txt 2>&1
python tools/dram_summary.py gpurun_out/r2v2_conv_dram.csv gpurun_out/r2v2_dram_per_launch.json "x" 2>&1 | tail -4
python tools/launch_summary.py gpurun_out/r2v2_launches_all.csv | head -14
