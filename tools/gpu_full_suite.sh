#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short > gpurun_out/k_all.log 2>&1; echo "kernels rc=$?"
tail -4 gpurun_out/k_all.log
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short > gpurun_out/m_all.log 2>&1; echo "models rc=$?"
tail -6 gpurun_out/m_all.log | cut -c1-300
VCG_BENCH_LAYERS=gpurun_out/layers_eager.txt timeout 900 python bench.py --steps 3 --warmup 2 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "bench_eager rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --graph 1 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench_graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-400
head -40 gpurun_out/layers_eager.txt
