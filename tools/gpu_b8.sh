#!/bin/bash
# per-GPU workload of the 8-GPU strong-scaling point (global batch 64 -> 8 images per GPU), on one GPU
mkdir -p gpurun_out
VCG_BENCH_LAYERS=gpurun_out/layers_b8.txt timeout 600 python bench.py --global-batch 8 --steps 5 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/bench_b8_eager.log 2>&1; echo "eager rc=$?"
timeout 600 python bench.py --global-batch 8 --steps 10 --warmup 3 --graph 1 --no-cpu-baseline > gpurun_out/bench_b8_graph.log 2>&1; echo "graph rc=$?"
tail -1 gpurun_out/bench_b8_graph.log | cut -c1-330
head -50 gpurun_out/layers_b8.txt
