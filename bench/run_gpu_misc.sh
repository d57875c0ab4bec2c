#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_engine_gpu.py -q --tb=short -m gpu > gpurun_out/engine_tests.log 2>&1
echo "== engine tests exit $?"; tail -n 6 gpurun_out/engine_tests.log | cut -c1-300
bash bench/sanitize.sh
rm -rf /tmp/synth_demo
timeout 600 python -m distributed_vgg_f_b200 -iu tcp://127.0.0.1:29999 -rn 0 -ws 1 -rd /tmp/synth_demo --synthetic 512 -ep 4 \
    -lr 0.00005 -mb 64 --profile events --log-jsonl gpurun_out/train_demo.jsonl > gpurun_out/train_demo.log 2>&1
echo "== VGG-F training demo exit $?"; grep -E "Epoch|Perf" gpurun_out/train_demo.log
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --model vgg16 --num-classes 1000 --batch 96 > gpurun_out/bench_vgg16_1000.log 2>&1
echo "== config#3 vgg16/1000/bs96 exit $?"; grep '^{' gpurun_out/bench_vgg16_1000.log | cut -c1-300
