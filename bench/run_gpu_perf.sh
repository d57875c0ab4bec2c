#!/bin/bash
mkdir -p gpurun_out

for grp in "gemm" "conv" "augment or bias_grad or maxpool or conv0"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -m gpu -k "$grp" > "gpurun_out/k_${name}.log" 2>&1
  echo "== $grp: exit $?"; tail -n 2 "gpurun_out/k_${name}.log"
done
timeout 300 python -m pytest tests/test_engine_gpu.py -q --tb=line -m gpu > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 4 gpurun_out/engine_tests.log | cut -c1-300
LB_TORCH=0 timeout 600 python bench/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "== layer bench exit $?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_native.log 2>&1
echo "== bench native exit $?"; tail -n 2 gpurun_out/bench_native.log | cut -c1-400
