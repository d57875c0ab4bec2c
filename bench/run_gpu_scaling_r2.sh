#!/bin/bash
# Round-2 scaling investigation on N GPUs of one box:
#   gpurun --gpus N --timeout 900 -- 'bash bench/run_gpu_scaling_r2.sh N [tests] [sweep] [ref]'
# 1. (tests) whole tests/test_multigpu.py incl. the experimental ZeRO-1 test
# 2. (sweep) all-reduce sweep: thin-CTA kernel, pack / wire-only entry, CTA counts 8..128, vs NCCL
# 3. bench.py at N=1 and N with the knobs that matter for overlap:
#      --comm-ctas {16,48,128}, B200_DYNAMIC_TILES {0,1}
N=${1:-2}
shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
port=29600
tr() { port=$((port+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
line() { grep '^{' "$1" | python -c 'import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d.get("impl"), "N=%s" % d.get("n_gpus"), d.get("ms_per_step"), "ms/step", d.get("value"), "img/s  e2e", (d.get("e2e") or {}).get("value"), "loss", d.get("final_loss"), "ar", d.get("allreduce"))' 2>&1 | cut -c1-400; }
for what in "$@"; do
  case $what in
    tests)
      B200_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu > gpurun_out/r2_multigpu_tests_n$N.log 2>&1
      echo "== multigpu tests exit $?"; tail -n 6 gpurun_out/r2_multigpu_tests_n$N.log | cut -c1-300 ;;
    sweep)
      AR_MAX_ELEMS=$((64*1024*1024)) tr bench/allreduce_sweep.py > gpurun_out/r2_allreduce_sweep_n$N.log 2>&1
      echo "== sweep exit $?"; grep -E "^n=" gpurun_out/r2_allreduce_sweep_n$N.log | cut -c1-200
      cp gpurun_out/allreduce_sweep_ws$N.json gpurun_out/r2_allreduce_sweep_ws$N.json 2>/dev/null ;;
    ref)
      tr bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_reference_n$N.log 2>&1
      echo "== reference N=$N: $(line gpurun_out/r2_bench_reference_n$N.log)" ;;
  esac
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/r2_bench_n1_on_box$N.log 2>&1
echo "== native N=1: $(line gpurun_out/r2_bench_n1_on_box$N.log)"
for dyn in 0 1; do
  for ctas in ${CTAS:-16 48 128}; do
    f=gpurun_out/r2_bench_n${N}_dyn${dyn}_c${ctas}.log
    B200_DYNAMIC_TILES=$dyn tr bench.py --gpus $N --steps 30 --warmup 5 --comm-ctas $ctas > $f 2>&1
    echo "== native N=$N dyn=$dyn ctas=$ctas: $(line $f)"
  done
done
