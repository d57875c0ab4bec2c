#!/bin/bash
# BASELINE config #5: the reference README's mini-batch table (mb in {16,32,64,96}, 1 vs 2 ranks) on
# synthetic data.  Usage: bash bench/mb_sweep.sh [impl]   (impl = native | reference)
IMPL=${1:-native}
mkdir -p gpurun_out
OUT=gpurun_out/mb_sweep_${IMPL}.jsonl
: > $OUT
for mb in 16 32 64 96; do
  python bench.py --impl $IMPL --gpus 1 --steps 10 --warmup 3 --batch $mb | grep '^{' >> $OUT
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 \
      bench.py --impl $IMPL --gpus 2 --steps 10 --warmup 3 --batch $mb | grep '^{' >> $OUT
done
python - <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mb_sweep_native.jsonl")]
print("| mb | 1 GPU img/s | 2 GPU img/s | speed-up |")
for i in range(0, len(rows), 2):
    a, b = rows[i], rows[i + 1]
    print("| %d | %.0f | %.0f | %.2fx |" % (a["config"]["per_gpu_batch"], a["value"], b["value"], b["value"] / a["value"]))
PY
