#!/bin/bash
mkdir -p gpurun_out
VCG_BENCH_LAYERS=gpurun_out/layers_eager.txt timeout 900 python bench.py --steps 3 --warmup 2 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "bench_eager rc=$?"
head -50 gpurun_out/layers_eager.txt
