# round-2 GPU job 20 (1 GPU): deferred weight-gradient join + high-priority capture stream
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_j20_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_j20_pytest.log)"
grep -E "^(FAILED|E  )" gpurun_out/r2_j20_pytest.log | head -20
for gb in 8 64; do
  timeout 300 python bench.py --global-batch $gb --steps 10 --warmup 3 --profile 0 --gpu-reference 0 --no-cpu-baseline > gpurun_out/r2_j20_gb$gb.log 2>&1; echo "gb$gb rc=$? $(grep '^{' gpurun_out/r2_j20_gb$gb.log | cut -c90-210)"
done
timeout 300 python tools/step_timeline.py --batch 8 --out gpurun_out/r2_j20_timeline_b8.csv > gpurun_out/r2_j20_timeline_b8.txt 2>&1; echo "b8 rc=$?"; cat gpurun_out/r2_j20_timeline_b8.txt | tail -20
