#!/bin/bash
# 2-GPU: data-parallel parity (fp32 mode) + scaling bench (graph and eager)
mkdir -p gpurun_out; rm -f gpurun_out/summary8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?" >> gpurun_out/summary8.txt
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --graph 1 > gpurun_out/bench_g2_graph.log 2>&1; echo "bench2_graph rc=$?" >> gpurun_out/summary8.txt
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --graph 0 > gpurun_out/bench_g2_eager.log 2>&1; echo "bench2_eager rc=$?" >> gpurun_out/summary8.txt
cat gpurun_out/summary8.txt; tail -3 gpurun_out/dp_check.log | cut -c1-400; tail -2 gpurun_out/bench_g2_graph.log | cut -c1-400; tail -2 gpurun_out/bench_g2_eager.log | cut -c1-400
