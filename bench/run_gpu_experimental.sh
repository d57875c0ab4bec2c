#!/bin/bash
# One gpurun call that validates the three experimental variants (docs/EXPERIMENTAL.md) against the default
# build and measures each: numerics first (a failing variant is skipped in the bench), then 1-GPU bench.py.
#   gpurun --timeout 900 -- 'bash bench/run_gpu_experimental.sh'
mkdir -p gpurun_out
run() { # name, env assignment(s), pytest -k expression
  local name=$1 envs=$2 expr=$3
  env $envs timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -m gpu -k "$expr" \
      > gpurun_out/exp_${name}_pytest.log 2>&1
  local rc=$?
  echo "== $name pytest exit $rc: $(tail -n 1 gpurun_out/exp_${name}_pytest.log | cut -c1-120)"
  return $rc
}
bench() { # name, env assignment(s)
  env $2 timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/exp_$1_bench.log 2>&1
  echo "== $1 bench: $(grep '^{' gpurun_out/exp_$1_bench.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], "ms/step", d["value"], "img/s, final loss", d.get("final_loss"))' 2>&1 | cut -c1-160)"
}
bench default "B200_NOP=1"
run pool "B200_EXPERIMENTAL=1" "fprop_pool_fused" && bench pool "B200_FUSE_POOL=1"
run tmaepi "B200_HALO_TMA_EPI=1" "conv_fprop or conv_dgrad or engine or backward or train" && bench tmaepi "B200_HALO_TMA_EPI=1"
run dyntiles "B200_DYNAMIC_TILES=1" "gemm or conv or engine or backward or train" && bench dyntiles "B200_DYNAMIC_TILES=1"
# the fused ZeRO-1 step needs >= 2 GPUs (gpurun --gpus 2 ...):
if [ "$(python -c 'import torch; print(torch.cuda.device_count())')" -ge 2 ]; then
  B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_multigpu.py -x -q -m gpu -k zero1 > gpurun_out/exp_zero1_pytest.log 2>&1
  echo "== zero1 pytest exit $?: $(tail -n 1 gpurun_out/exp_zero1_pytest.log | cut -c1-120)"
  N=$(python -c 'import torch; print(torch.cuda.device_count())')
  for z in "" "--zero1"; do
    timeout 300 python bench.py --gpus $N --steps 30 --warmup 5 $z > gpurun_out/exp_zero1_bench${z:+_on}.log 2>&1
    echo "== N=$N bench ${z:-default}: $(grep '^{' gpurun_out/exp_zero1_bench${z:+_on}.log | cut -c1-200)"
  done
fi
