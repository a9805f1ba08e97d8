#!/bin/bash
# round-2 first visit: MMA rate for N=144.., launch lists of the covariance step, ncu of the
# second-moment kernels and the insert loop
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
tools/cuda/build/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1; echo "mma_rate $?"; grep "grid=148 distinct=2" gpurun_out/r2_mma_rate.txt
for b in 32 128; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cov_launches_b$b.csv python tools/prof_r2.py cov $b 3 > gpurun_out/r2_cov_b$b.log 2>&1; echo "cov launches b$b $?"
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gram_tc -s 2 -c 1 -f -o gpurun_out/r2_gram python tools/prof_r2.py cov 32 3 > gpurun_out/r2_ncu_gram.log 2>&1; echo "gram $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:reduce_partials -s 2 -c 1 -f -o gpurun_out/r2_reduce python tools/prof_r2.py cov 32 3 > gpurun_out/r2_ncu_reduce.log 2>&1; echo "reduce $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:insert_loop -s 1 -c 1 -f -o gpurun_out/r2_insert python tools/prof_r2.py insert 200 > gpurun_out/r2_ncu_insert.log 2>&1; echo "insert $?"
ls -la gpurun_out/*.ncu-rep | tail -5
