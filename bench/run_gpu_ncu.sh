#!/bin/bash
mkdir -p gpurun_out
# (1) every launch of one training step with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 260 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 4 > gpurun_out/ncu_launches.log 2>&1
echo "== launch list exit $?"; wc -l gpurun_out/launches.csv
# (2) full-set capture of the representative kernels
ncu --set full --clock-control none --import-source on -k regex:"umma_kernel|conv_halo|wgrad_halo|conv0_kernel" -o gpurun_out/prof_conv_v4 \
    python bench/ncu_conv.py > gpurun_out/ncu_conv_v4.log 2>&1
echo "== ncu full exit $?"; ls -la gpurun_out/prof_conv_v4.ncu-rep
