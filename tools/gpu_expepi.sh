#!/bin/bash
mkdir -p gpurun_out
L="e0_b64 dU4_b64 dU3_b64 p_c64k3 p_c64k1"
for e in 0 1 2 4; do
echo "exp=$e"
VCG_NO_FOLD=1 VCG_EXP_EPI=$e timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-100
done
