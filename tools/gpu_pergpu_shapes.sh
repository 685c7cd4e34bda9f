#!/bin/bash
# the per-GPU workloads of the 2/4/8-GPU strong-scaling points (global batch 64) on one GPU: no crash, finite metrics
mkdir -p gpurun_out
for b in 32 16 8; do
  timeout 300 python bench.py --global-batch $b --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith(chr(123)):
        d=json.loads(l); print('B=$b', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), d['final_metrics'])"
done
