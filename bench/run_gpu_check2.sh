#!/bin/bash
# 2-GPU check after a comm change: zero1/equivalence tests, N=1 and N=2 bench (zero1 off/on), timeline
N=2; mkdir -p gpurun_out; port=31200
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print("N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", round(d.get("final_loss"),4), "ar", (d.get("allreduce") or {}).get("ms_per_step"), d["config"]["parallelism"])' 2>&1 | cut -c1-300; }
timeout 600 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -k "zero1 or equivalence or stress" > gpurun_out/check2_tests.log 2>&1
echo "== tests exit $?: $(tail -n 1 gpurun_out/check2_tests.log)"
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/check2_n1.log 2>&1; echo "== N=1: $(line gpurun_out/check2_n1.log)"
for z in off on; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --zero1 $z > gpurun_out/check2_n2_z$z.log 2>&1
  echo "== N=2 zero1=$z: $(line gpurun_out/check2_n2_z$z.log)"
done
port=$((port+1))
B200_SPLIT_COMM=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --zero1 off > gpurun_out/check2_n2_onestream.log 2>&1
echo "== N=2 zero1=off one side stream: $(line gpurun_out/check2_n2_onestream.log)"
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench/step_timeline.py --zero1 off --tag _check2 > gpurun_out/timeline_n2_check2.txt 2>&1
tail -n 1 gpurun_out/timeline_n2_check2.txt
