#!/bin/bash
# 1-GPU validation: the whole GPU test suite, the bench, and the README-table protocol through the CLI path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q --tb=short -m gpu > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r2a_pytest_gpu.log)"
grep -E "^\{" gpurun_out/r2a_pytest_gpu.log | head -3
