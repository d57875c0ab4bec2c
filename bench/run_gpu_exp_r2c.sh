#!/bin/bash
# after the evict-first change: optimizer CTAs per SM x {N=1, N} ; zero1 ; timelines
N=${1:-2}
mkdir -p gpurun_out
port=29900
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print("N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", round(d.get("final_loss"),4), "ar", (d.get("allreduce") or {}).get("ms_per_step"), d["config"]["parallelism"])' 2>&1 | cut -c1-300; }
for a in 8 4 1; do
  B200_ADAM_CTAS_PER_SM=$a timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/r2c_n1_adam$a.log 2>&1
  echo "== N=1 adam_ctas/sm=$a: $(line gpurun_out/r2c_n1_adam$a.log)"
done
run() { # tag, env assignments..., --, bench args
  local tag=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  port=$((port+1))
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2c_n${N}_$tag.log 2>&1
  echo "== N=$N $tag: $(line gpurun_out/r2c_n${N}_$tag.log)"
}
run adam8 B200_ADAM_CTAS_PER_SM=8 --
run adam4 B200_ADAM_CTAS_PER_SM=4 --
run adam1 B200_ADAM_CTAS_PER_SM=1 --
run adam4_c16 B200_ADAM_CTAS_PER_SM=4 -- --comm-ctas 16
run adam4_c128 B200_ADAM_CTAS_PER_SM=4 -- --comm-ctas 128
run zero1 B200_ADAM_CTAS_PER_SM=4 -- --zero1
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench/step_timeline.py --tag _ef > gpurun_out/timeline_n${N}_ef.txt 2>&1
timeout 300 python bench/step_timeline.py --tag _ef > gpurun_out/timeline_n1_ef.txt 2>&1
tail -n 1 gpurun_out/timeline_n${N}_ef.txt gpurun_out/timeline_n1_ef.txt
