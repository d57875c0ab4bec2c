#!/bin/bash
# One 8-GPU lease: correctness at ws=8, then the scaling numbers and their explanation.
#   gpurun --gpus 8 --timeout 1500 -- 'bash bench/run_gpu_n8_r2.sh'
N=${1:-8}
mkdir -p gpurun_out
port=30100
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print("N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", round(d.get("final_loss"),4), "ar", (d.get("allreduce") or {}), d["config"]["parallelism"])' 2>&1 | cut -c1-420; }
B200_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu > gpurun_out/r2_multigpu_tests_n$N.log 2>&1
echo "== multigpu tests exit $?: $(tail -n 1 gpurun_out/r2_multigpu_tests_n$N.log)"
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/r2n8_n1.log 2>&1
echo "== N=1: $(line gpurun_out/r2n8_n1.log)"
run() { # tag, env assignments..., --, bench args
  local tag=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  port=$((port+1))
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2n8_n${N}_$tag.log 2>&1
  echo "== N=$N $tag: $(line gpurun_out/r2n8_n${N}_$tag.log)"
}
run default X=1 --
run zero1 X=1 -- --zero1
run c128 X=1 -- --comm-ctas 128
run c16 X=1 -- --comm-ctas 16
run zero1_c256 B200_ZERO1_CTAS=128 X=1 -- --zero1 --comm-ctas 128
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench/step_timeline.py --tag _r2 > gpurun_out/timeline_n${N}_r2.txt 2>&1
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench/step_timeline.py --zero1 --tag _r2_zero1 > gpurun_out/timeline_n${N}_r2_zero1.txt 2>&1
tail -n 1 gpurun_out/timeline_n${N}_r2.txt gpurun_out/timeline_n${N}_r2_zero1.txt
port=$((port+1))
AR_MAX_ELEMS=$((64*1024*1024)) AR_CTAS=16,48,128 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench/allreduce_sweep.py > gpurun_out/r2_allreduce_sweep_n$N.log 2>&1
echo "== sweep exit $?"; grep -E "^n=" gpurun_out/r2_allreduce_sweep_n$N.log | cut -c1-160
cp gpurun_out/allreduce_sweep_ws$N.json gpurun_out/r2_allreduce_sweep_ws$N.json 2>/dev/null
