#!/bin/bash
# First-contact GPU check: every kernel family in its own process (a device trap in one family
# must not poison the others).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "gemm_kk" "gemm_mnk" "gemm_mnmn" "gemm_bf16" "conv_fprop" "conv_dgrad" "conv_wgrad" \
           "maxpool or avgpool or bias_grad or fc_epilogues" "cross_entropy or adam or sgd" "augment"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -m gpu -k "$grp" \
      > "gpurun_out/k_${name}.log" 2>&1
  echo "== $grp: exit $?"; tail -n 3 "gpurun_out/k_${name}.log"
done
