#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary10.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short > gpurun_out/k_all.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary10.txt
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -s -k "bf16 and (vae_forward or discriminator or cyclevaegan)" > gpurun_out/m_bf16.log 2>&1; echo "models rc=$?" >> gpurun_out/summary10.txt
timeout 300 python tools/bench_conv.py eR_b64 eD2 > gpurun_out/conv_normal.log 2>&1
VCG_EXP_MN=1 timeout 300 python tools/bench_conv.py eR_b64 eD2 > gpurun_out/conv_mn.log 2>&1
VCG_BENCH_LAYERS=gpurun_out/layers_eager.txt timeout 900 python bench.py --steps 3 --warmup 2 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "bench_eager rc=$?" >> gpurun_out/summary10.txt
timeout 900 python bench.py --steps 5 --warmup 3 --graph 1 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench_graph rc=$?" >> gpurun_out/summary10.txt
cat gpurun_out/summary10.txt; tail -3 gpurun_out/k_all.log; grep -E "passed|failed|^E   Assert" gpurun_out/m_bf16.log | cut -c1-250 | tail -5; cat gpurun_out/conv_normal.log gpurun_out/conv_mn.log
