#!/bin/bash
# compute-sanitizer passes over the kernel-level GPU tests (the reference has no
# race / memory checking).  racecheck covers the shared-memory protocols (mbarrier rings, the
# epilogue mailboxes, transposes); memcheck the global accesses of every kernel.  Logs go to
# gpurun_out/ and their summaries are kept under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="tests/test_gpu_fastpath_kernels.py tests/test_gpu_kernels.py"
[ -f tests/test_gpu_kernels.py ] || T="tests/test_gpu_fastpath_kernels.py"
SEL='modconv_up_fused or rgb_combine or pixel_norm or demod_multi or second_moment or styled_conv_forward'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 --report-api-errors no \
    python -m pytest $T tests/test_gpu_parity.py -q -x --timeout 1100 -p no:cacheprovider -k "$SEL" \
    > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/r2_sanitize_$tool.log | head -12
done
