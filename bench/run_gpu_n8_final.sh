#!/bin/bash
# One 8-GPU lease, final code: (1) the ws=8-relevant multi-GPU tests, (2) bench at N=4 and N=8, (3) the N=8 step timeline
# with the per-rank skew.  The 2-rank tests of tests/test_multigpu.py run on a 2-GPU lease (bench/run_gpu_check2.sh).
N=8; mkdir -p gpurun_out; port=31500
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print("N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", round(d.get("final_loss"),4), "ar", (d.get("allreduce") or {}), d["config"]["parallelism"])' 2>&1 | cut -c1-460; }
timeout 600 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -k "matches_nccl or wide or stress or fake_nodes" > gpurun_out/r2_multigpu_tests_ws8_final.log 2>&1
echo "== ws=8 tests exit $?: $(tail -n 1 gpurun_out/r2_multigpu_tests_ws8_final.log)"
for n in 8 4; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2_final_bench_n$n.log 2>&1
  echo "== N=$n: $(line gpurun_out/r2_final_bench_n$n.log)"
done
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench/step_timeline.py --tag _final > gpurun_out/timeline_n8_final.txt 2>&1
grep -E "per-rank|exposed" gpurun_out/timeline_n8_final.txt
timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/r2_final_bench_n1_on_box8.log 2>&1
echo "== N=1 (GPU 0 of the same box): $(line gpurun_out/r2_final_bench_n1_on_box8.log)"
