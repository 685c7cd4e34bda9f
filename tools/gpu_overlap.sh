#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -x > gpurun_out/m_overlap.log 2>&1; echo "models rc=$?"; tail -3 gpurun_out/m_overlap.log | cut -c1-200
for b in 64 8; do
  timeout 300 python bench.py --global-batch $b --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith(chr(123)):
        d=json.loads(l); print('B=$b', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), d['final_metrics'])"
done
