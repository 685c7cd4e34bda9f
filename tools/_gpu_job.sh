# round-2 GPU job 27 (1 GPU): final validation -- full GPU suite, default bench line, smoke
timeout 150 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/r2_j27_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_j27_pytest.log)"
grep -E "^(FAILED|E  )" gpurun_out/r2_j27_pytest.log | head -10
timeout 100 python bench.py > gpurun_out/r2_j27_default.log 2>&1; echo "bench rc=$? $(grep '^{' gpurun_out/r2_j27_default.log | cut -c90-220)"
timeout 60 python __graft_entry__.py smoke > gpurun_out/r2_j27_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2_j27_smoke.log | cut -c1-200)"
