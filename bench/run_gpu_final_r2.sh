#!/bin/bash
# Final validation of the tree on one B200: smoke(), the whole GPU test suite, bench.py with its defaults.
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "== smoke exit $?: $(tail -n 1 gpurun_out/r2_final_smoke.log)"
timeout 600 python -m pytest tests -q --tb=short -m gpu > gpurun_out/r2_final_pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?: $(tail -n 1 gpurun_out/r2_final_pytest_gpu.log)"
timeout 200 python bench.py > gpurun_out/r2_final_bench_n1.log 2>&1
grep '^{' gpurun_out/r2_final_bench_n1.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("== bench (defaults)", d["ms_per_step"], "ms", d["value"], "img/s e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]/d["steps"], "loss", d["final_loss"], d["clocks"])'
