#!/bin/bash
# compute-sanitizer passes over the kernel library.  Runs on ONE GPU; the tcgen05 /
# TMA kernels are checked with memcheck + synccheck, the element-wise kernels additionally with
# racecheck (racecheck does not model async-proxy (TMA / tcgen05) shared-memory traffic).
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
K='gemm_kk and 128-64-64 or conv_fprop and 2-16-16 or conv_wgrad and 2-16-16 or conv_dgrad and 2-32-32 or maxpool or cross_entropy and 64-3 or adam or bias_grad and 64-3 or head_ce or fprop_pool_fused'
for tool in memcheck synccheck; do
  timeout 900 $SAN --tool $tool --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "$K" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "== compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
timeout 900 $SAN --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu \
    -k "maxpool or cross_entropy and 64-3 or bias_grad and 64-3 or adam or head_ce" > gpurun_out/sanitize_racecheck.log 2>&1
echo "== compute-sanitizer racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
