# round-2 GPU job: full GPU suite, then the per-GPU-batch sweep of the step (lanes on / off), event overhead check
python -m pytest tests -m gpu -q -x --deselect tests/test_baseline_configs_gpu.py > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -4 gpurun_out/r2_t3.log
python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -s > gpurun_out/r2_t3_baseline.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3_baseline.log
grep -E "^\[config|passed|failed|Error|assert" gpurun_out/r2_t3_baseline.log | cut -c1-400 | tail -30
for gb in 64 32 16 8; do for ln in 0 1; do python bench.py --global-batch $gb --lanes $ln --steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 > gpurun_out/r2_b3_gb${gb}_l${ln}.log 2>&1; echo "gb$gb lanes$ln $(tail -1 gpurun_out/r2_b3_gb${gb}_l${ln}.log | cut -c90-200)"; done; done
python bench.py --global-batch 64 --profile 0 --steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 > gpurun_out/r2_b3_gb64_noprof.log 2>&1; echo "gb64 noprofile $(tail -1 gpurun_out/r2_b3_gb64_noprof.log | cut -c90-200)"
python bench.py --global-batch 8 --profile 0 --steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 > gpurun_out/r2_b3_gb8_noprof.log 2>&1; echo "gb8 noprofile $(tail -1 gpurun_out/r2_b3_gb8_noprof.log | cut -c90-200)"
VCG_BENCH_LAYERS=gpurun_out/r2_layers_b8_graph.txt python bench.py --global-batch 8 --lanes 0 --steps 5 --warmup 3 --no-cpu-baseline --gpu-reference 0 > gpurun_out/r2_b3_gb8_layers.log 2>&1
