#!/bin/bash
# 2-GPU data-parallel parity (gradients / weights vs a single-GPU run of the same global batch) + 2-GPU bench
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_check.py > gpurun_out/dp_check_2gpu.log 2>&1; echo "dp_check rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|Warning\|warn" gpurun_out/dp_check_2gpu.log | tail -8 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g2_graph.log 2>&1; echo "g2 rc=$?"
grep '"metric"' gpurun_out/bench_g2_graph.log | cut -c1-330
