#!/bin/bash
mkdir -p gpurun_out
for st in 8 7 6 4 3; do echo "stages $st"; VCG_FOLD_STAGES=$st timeout 300 python tools/bench_conv.py d5_b64 2>&1 | tail -1; done
