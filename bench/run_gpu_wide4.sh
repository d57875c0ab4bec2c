mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu.py -q --tb=short -m gpu -k "wide" > gpurun_out/r2_multigpu_tests_ws4_wide.log 2>&1
echo "== exit $?: $(tail -n 1 gpurun_out/r2_multigpu_tests_ws4_wide.log)"
