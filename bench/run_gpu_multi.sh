#!/bin/bash
# usage: run_gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -x > gpurun_out/multigpu_tests.log 2>&1
echo "== multigpu tests exit $?"; tail -n 30 gpurun_out/multigpu_tests.log
for impl in native reference; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --impl $impl --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${impl}_n$N.log 2>&1
  echo "== bench $impl N=$N exit $?"; tail -n 3 gpurun_out/bench_${impl}_n$N.log | cut -c1-1500
done
