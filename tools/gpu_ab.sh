#!/bin/bash
# same-box A/B of the graph-replayed step: does time saved in memory-bound passes show up under the power cap?
mkdir -p gpurun_out
for i in 1 2; do
for cfg in "default" "VCG_XF_WAVES=8" "VCG_WTC2=0"; do
  if [ "$cfg" = "default" ]; then e=""; else e="$cfg"; fi
  env $e timeout 300 python bench.py --steps 8 --warmup 3 --graph 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$cfg', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
done
