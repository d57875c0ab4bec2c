#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 25 gpurun_out/engine_tests.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_native.log 2>&1
echo "== bench native exit $?"; tail -n 5 gpurun_out/bench_native.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1
echo "== bench reference exit $?"; tail -n 5 gpurun_out/bench_ref.log
