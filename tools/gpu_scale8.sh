#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g8_graph.log 2>&1; echo "g8 graph rc=$?"
tail -1 gpurun_out/bench_g8_graph.log | cut -c1-300
VCG_BENCH_LAYERS=gpurun_out/layers_g8_eager.txt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 2 --graph 0 --no-cpu-baseline > gpurun_out/bench_g8_eager.log 2>&1; echo "g8 eager rc=$?"
tail -1 gpurun_out/bench_g8_eager.log | cut -c1-300
head -30 gpurun_out/layers_g8_eager.txt
