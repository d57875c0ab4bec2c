#!/bin/bash
# 1 GPU: BASELINE configs #2 (context arm, sustained), #3 (VGG-16/1000, mb 96), #5 (reference arm), ncu of the non-conv kernels
mkdir -p gpurun_out
J() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d.get("impl"), d["config"].get("model"), "mb", d["config"].get("per_gpu_batch"), "steps", d.get("steps"), "|", d.get("ms_per_step"), "ms/step", d.get("value"), "img/s | e2e", (d.get("e2e") or {}).get("value"), "| clocks", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))' 2>&1 | cut -c1-260; }
bash bench/run_gpu_ncu_r2.sh
timeout 200 python bench.py --impl reference-bf16 --steps 20 --warmup 5 > gpurun_out/r2b_refbf16_n1.log 2>&1; echo "== $(J gpurun_out/r2b_refbf16_n1.log)"
timeout 200 python bench.py --steps 1000 --warmup 20 > gpurun_out/r2b_native_sustained_n1.log 2>&1; echo "== sustained $(J gpurun_out/r2b_native_sustained_n1.log)"
timeout 200 python bench.py --impl reference --steps 300 --warmup 10 > gpurun_out/r2b_reference_sustained_n1.log 2>&1; echo "== sustained $(J gpurun_out/r2b_reference_sustained_n1.log)"
timeout 200 python bench.py --model vgg16 --num-classes 1000 --batch 96 --steps 20 --warmup 5 > gpurun_out/r2b_vgg16_native_n1.log 2>&1; echo "== $(J gpurun_out/r2b_vgg16_native_n1.log)"
timeout 300 python bench.py --impl reference --model vgg16 --num-classes 1000 --batch 96 --steps 10 --warmup 3 > gpurun_out/r2b_vgg16_reference_n1.log 2>&1; echo "== $(J gpurun_out/r2b_vgg16_reference_n1.log)"
timeout 300 python bench.py --impl reference-bf16 --model vgg16 --num-classes 1000 --batch 96 --steps 10 --warmup 3 > gpurun_out/r2b_vgg16_refbf16_n1.log 2>&1; echo "== $(J gpurun_out/r2b_vgg16_refbf16_n1.log)"
rm -f gpurun_out/readme_table_reference.jsonl
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench/readme_table.py --impl reference --train-per-class 480 > gpurun_out/r2b_readme_reference_n1.log 2>&1
echo "== readme table reference N=1 exit $?"; grep '^{' gpurun_out/r2b_readme_reference_n1.log | cut -c1-220
