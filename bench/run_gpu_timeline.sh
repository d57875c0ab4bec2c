#!/bin/bash
# step timelines at 1 GPU and N GPUs:  gpurun --gpus N -- 'bash bench/run_gpu_timeline.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python bench/step_timeline.py > gpurun_out/timeline_n1.txt 2>&1; echo "== n1 exit $?"
for c in ${CTAS:-48}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
   bench/step_timeline.py --comm-ctas $c --tag _c$c > gpurun_out/timeline_n${N}_c$c.txt 2>&1; echo "== n$N c$c exit $?"
done
tail -n 3 gpurun_out/timeline_n*.txt
