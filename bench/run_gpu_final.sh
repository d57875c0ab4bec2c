#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; tail -n 3 gpurun_out/final_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "== smoke exit $?"; tail -n 2 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_reference_n1.log 2>&1
echo "== bench reference exit $?"; grep '^{' gpurun_out/final_bench_reference_n1.log | cut -c1-250
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_native_n1.log 2>&1
echo "== bench native exit $?"; grep '^{' gpurun_out/final_bench_native_n1.log | cut -c1-1800
