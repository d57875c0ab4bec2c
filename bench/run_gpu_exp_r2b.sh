#!/bin/bash
# comm-stream interference experiments at N GPUs (and N=1): thin optimizer CTAs, fused ZeRO-1, stream priority
N=${1:-2}
mkdir -p gpurun_out
port=29800
tr() { port=$((port+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print("N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", round(d.get("final_loss"),4), "ar", (d.get("allreduce") or {}).get("ms_per_step"), d["config"]["parallelism"])' 2>&1 | cut -c1-300; }
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -k "zero1 or equivalence" > gpurun_out/r2b_tests_n$N.log 2>&1
echo "== tests exit $?: $(tail -n 1 gpurun_out/r2b_tests_n$N.log)"
for a in 1 8; do
  B200_ADAM_CTAS_PER_SM=$a timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/r2b_n1_adam$a.log 2>&1
  echo "== N=1 adam_ctas/sm=$a: $(line gpurun_out/r2b_n1_adam$a.log)"
done
run() { # tag, env assignments..., --, bench args
  local tag=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  port=$((port+1))
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2b_n${N}_$tag.log 2>&1
  echo "== N=$N $tag: $(line gpurun_out/r2b_n${N}_$tag.log)"
}
run adam1 B200_ADAM_CTAS_PER_SM=1 --
run adam8 B200_ADAM_CTAS_PER_SM=8 --
run adam2 B200_ADAM_CTAS_PER_SM=2 --
run zero1 B200_ADAM_CTAS_PER_SM=1 -- --zero1
run zero1_c64 B200_ADAM_CTAS_PER_SM=1 B200_ZERO1_CTAS=64 -- --zero1
run zero1_prio0 B200_ADAM_CTAS_PER_SM=1 B200_COMM_PRIORITY=0 -- --zero1
run prio0 B200_ADAM_CTAS_PER_SM=1 B200_COMM_PRIORITY=0 --
run zero1_dyn B200_ADAM_CTAS_PER_SM=1 B200_DYNAMIC_TILES=1 -- --zero1
tr bench/step_timeline.py --zero1 --tag _zero1 > gpurun_out/timeline_n${N}_zero1.txt 2>&1
tr bench/step_timeline.py --tag _adam1 > gpurun_out/timeline_n${N}_adam1.txt 2>&1
timeout 300 python bench/step_timeline.py --tag _adam1 > gpurun_out/timeline_n1_adam1.txt 2>&1
tail -n 2 gpurun_out/timeline_n${N}_zero1.txt gpurun_out/timeline_n${N}_adam1.txt gpurun_out/timeline_n1_adam1.txt
