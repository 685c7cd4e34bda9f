#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "xform or conv_layer" > gpurun_out/k_xform.log 2>&1; echo "kernels rc=$?"
tail -15 gpurun_out/k_xform.log
for v in 0 1 2 3 4; do
VCG_XB_VARIANT=$v timeout 300 python tools/bench_xform.py > gpurun_out/bench_xform_v$v.log 2>&1; echo "bench_xform v$v rc=$?"
grep -E "gath|fold|total" gpurun_out/bench_xform_v$v.log
done
