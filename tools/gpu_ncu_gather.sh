#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xform_bwd_gather --launch-skip 0 -c 1 -f -o gpurun_out/r01g_gather_c64 python tools/bench_xform.py > gpurun_out/ncu_gather.log 2>&1; echo "ncu rc=$?"
