#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "xform or conv_layer" > gpurun_out/k_xform.log 2>&1; echo "kernels rc=$?"
tail -5 gpurun_out/k_xform.log
timeout 300 python tools/bench_xform.py > gpurun_out/bench_xform.log 2>&1; echo "bench_xform rc=$?"
cat gpurun_out/bench_xform.log
