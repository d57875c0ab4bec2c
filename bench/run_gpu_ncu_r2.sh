#!/bin/bash
# ncu --set full of everything in a training step that is NOT a convolution (those: profiles/r1_ncu_*):
# optimizer, pooling, input transform, first conv, loss, FC GEMMs and their epilogues.  1 GPU.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"adam_kernel|maxpool|unpool|augment|conv0_kernel|cross_entropy|fc_bias_act|fc_grad_act|bias_grad|GemmPolicy" \
    -o gpurun_out/prof_step_r2 -f python bench/ncu_step.py > gpurun_out/ncu_step_r2.log 2>&1
echo "== ncu exit $?"; ls -la gpurun_out/prof_step_r2.ncu-rep
