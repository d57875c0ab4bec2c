mkdir -p gpurun_out
K=none,probe_sleep_1,probe_ldst_1,probe_mufu_1,adam,d2d_engine
for c in 0 1; do
echo "=== carveout preference of the background kernels: $c"
KINDS=$K TAG=_carve$c B200_PROBE_CARVEOUT=$c B200_COMM_CARVEOUT=$c B200_ADAM_CTAS_PER_SM=4 timeout 300 python bench/interference.py 2>&1 | grep -v Warn | tee gpurun_out/interference_carve$c.txt
done
for a in 1 4 8; do
B200_ADAM_CTAS_PER_SM=$a timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | grep '^{' | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("adam/sm='$a' carve=1", d["ms_per_step"], d["value"])'
done
B200_COMM_CARVEOUT=0 B200_ADAM_CTAS_PER_SM=8 timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | grep '^{' | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("adam/sm=8 carve=0", d["ms_per_step"], d["value"])'
