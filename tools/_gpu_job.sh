# round-2 GPU job 7: suite, step at batch 64 / 8 after the epilogue-register and dhead_prepare fixes, unfolded-halo A/B
python -m pytest tests -m gpu -q > gpurun_out/r2_t7.log 2>&1; echo "suite rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_t7.log | cut -c1-200 | tail -12
grep -E "^E  " gpurun_out/r2_t7.log | cut -c1-250 | head -12
grep -E "\[config|\[white" gpurun_out/r2_t7.log | cut -c1-300
B="--steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 --profile 0"
for gb in 64 8; do for uf in 0 1024 4096; do python bench.py --global-batch $gb --unfolded-max-hw $uf $B > gpurun_out/r2_b7_gb${gb}_uf${uf}.log 2>&1; echo "gb$gb unfolded$uf $(tail -1 gpurun_out/r2_b7_gb${gb}_uf${uf}.log | cut -c90-200)"; done; done
python bench.py --global-batch 16 $B > gpurun_out/r2_b7_gb16.log 2>&1; echo "gb16 $(tail -1 gpurun_out/r2_b7_gb16.log | cut -c90-200)"
python bench.py --global-batch 32 $B > gpurun_out/r2_b7_gb32.log 2>&1; echo "gb32 $(tail -1 gpurun_out/r2_b7_gb32.log | cut -c90-200)"
