#!/bin/bash
mkdir -p gpurun_out
L="eR_b64 eD1_b64 dU2_b64 dU1_b64 dU3_b64 eD2 eD3"
timeout 300 python tools/bench_conv.py $L > gpurun_out/skipa_off.jsonl 2>&1
VCG_EXP_SKIPA=1 timeout 300 python tools/bench_conv.py $L > gpurun_out/skipa_on.jsonl 2>&1
paste -d'\n' gpurun_out/skipa_off.jsonl gpurun_out/skipa_on.jsonl | cut -c1-150
