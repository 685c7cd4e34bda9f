python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
for gb in 64 8; do for ln in 0 1; do python bench.py --global-batch $gb --lanes $ln --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_gb${gb}_l${ln}.log 2>&1; done; done
tail -3 gpurun_out/r2_t1.log
for f in gpurun_out/r2_b_gb*; do echo $f; tail -1 $f | cut -c1-200; done
