# round-2 GPU job 23 (1 GPU): same-box A/B of the step-schedule changes (A old, B deferred join + priority, C + small last D bucket, D small last bucket everywhere)
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -q -x --timeout 600 -k "lanes or graph or golden or losses" > gpurun_out/r2_j23_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_j23_pytest.log)"
grep -E "^(FAILED|E  )" gpurun_out/r2_j23_pytest.log | head -20
run() { # name batch flags
  timeout 300 python bench.py --global-batch $2 --steps 10 --warmup 3 --profile 0 --gpu-reference 0 --no-cpu-baseline $3 > gpurun_out/r2_j23_$1_gb$2.log 2>&1
  echo "$1 gb$2 rc=$? $(grep '^{' gpurun_out/r2_j23_$1_gb$2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3))")"
}
for rep in 1 2; do for gb in 8 64; do
  run A $gb "--wgrad-defer 0 --capture-priority 0 --tail-params 0"
  run B $gb "--tail-params 0"
  run C $gb ""
  run D $gb "--tail-params 1048576"
done; done
