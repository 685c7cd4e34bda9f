#!/bin/bash
# first GPU contact: kernel tests in risk order, then the conv micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -k "not conv_layer" > gpurun_out/t1_misc.log 2>&1; echo "misc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -k "conv_layer and fp32" > gpurun_out/t2_conv_fp32.log 2>&1; echo "conv_fp32 rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -k "conv_layer and bf16" > gpurun_out/t3_conv_bf16.log 2>&1; echo "conv_bf16 rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_misc.log gpurun_out/t2_conv_fp32.log gpurun_out/t3_conv_bf16.log
cat gpurun_out/bench_conv.log | tail -12
