#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q --tb=short -k "xform or norm" > gpurun_out/xf_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/xf_tests.log
timeout 300 python tools/bench_xform.py 2>&1 | grep "^fwd\|total"
