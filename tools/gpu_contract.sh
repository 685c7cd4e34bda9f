#!/bin/bash
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -4 gpurun_out/bench_default.log | cut -c1-600
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -4 gpurun_out/bench_ref.log | cut -c1-700
