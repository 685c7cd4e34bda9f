# round-2 GPU job 16 (1 GPU): full GPU suite, smoke, default bench line
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_final_pytest.log)"
grep -E "^(FAILED|E  )" gpurun_out/r2_final_pytest.log | head -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 gpurun_out/r2_final_smoke.log | cut -c1-200)"
timeout 600 python bench.py > gpurun_out/r2_final_default.log 2>&1; echo "bench rc=$? $(grep '^{' gpurun_out/r2_final_default.log | cut -c90-220)"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_refarm.log 2>&1; echo "ref arm rc=$? $(grep '^{' gpurun_out/r2_final_refarm.log | cut -c1-200)"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
