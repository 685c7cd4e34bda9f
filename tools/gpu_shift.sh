#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
VCG_TC2=0 VCG_EXP_SHIFT=$m timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=line -k "conv_layer and bf16" > gpurun_out/shift_$m.log 2>&1
echo "mode $m rc=$?"; tail -25 gpurun_out/shift_$m.log | cut -c1-220
done
