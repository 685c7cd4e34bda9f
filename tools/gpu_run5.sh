#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary5.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -k "d5 or adam or reparam or dhead" > gpurun_out/k1.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary5.txt
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -s > gpurun_out/m_all.log 2>&1; echo "models rc=$?" >> gpurun_out/summary5.txt
VCG_BENCH_LAYERS=gpurun_out/layers_eager.txt timeout 900 python bench.py --steps 3 --warmup 2 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "bench_eager rc=$?" >> gpurun_out/summary5.txt
timeout 900 python bench.py --steps 5 --warmup 3 --graph 1 > gpurun_out/bench_graph.log 2>&1; echo "bench_graph rc=$?" >> gpurun_out/summary5.txt
cat gpurun_out/summary5.txt; tail -3 gpurun_out/k1.log; grep -E "^\[|passed|failed|^E   Assert" gpurun_out/m_all.log | cut -c1-250 | tail -40
tail -2 gpurun_out/bench_eager.log | cut -c1-1500; head -30 gpurun_out/layers_eager.txt; tail -4 gpurun_out/bench_graph.log | cut -c1-2500
