#!/bin/bash
# usage: run_gpu_multi.sh N [tests]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -x > gpurun_out/multigpu_tests_n$N.log 2>&1
  echo "== multigpu tests exit $?"; tail -n 8 gpurun_out/multigpu_tests_n$N.log
fi
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench/allreduce_sweep.py > gpurun_out/allreduce_sweep_n$N.log 2>&1
echo "== sweep exit $?"; grep -E "^n=" gpurun_out/allreduce_sweep_n$N.log | cut -c1-260; grep -ciE "NVLS" gpurun_out/allreduce_sweep_n$N.log
for impl in native reference; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --impl $impl --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${impl}_n$N.log 2>&1
  echo "== bench $impl N=$N exit $?"; grep '^{' gpurun_out/bench_${impl}_n$N.log | cut -c1-330
done
