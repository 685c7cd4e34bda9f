# round-2 GPU job 14: several output rows per MMA in the 7x7 fold kernel
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv_layer" > gpurun_out/r2_t14.log 2>&1; echo "conv kernel tests rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_t14.log | cut -c1-200 | tail -6
grep -E "^E  " gpurun_out/r2_t14.log | cut -c1-250 | head -8
B="--steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 --profile 0"
timeout 300 python bench.py --global-batch 64 $B > gpurun_out/r2_b14_gb64.log 2>&1; echo "gb64 $(tail -1 gpurun_out/r2_b14_gb64.log | cut -c90-200)"
timeout 300 python bench.py --global-batch 8 $B > gpurun_out/r2_b14_gb8.log 2>&1; echo "gb8 $(tail -1 gpurun_out/r2_b14_gb8.log | cut -c90-200)"
VCG_BENCH_LAYERS=gpurun_out/r2_layers_b64_v2.txt timeout 300 python bench.py --global-batch 64 --steps 5 --warmup 3 --no-cpu-baseline --gpu-reference 0 > gpurun_out/r2_b14_prof.log 2>&1; grep -E "fold|k7" gpurun_out/r2_layers_b64_v2.txt | cut -c1-150
