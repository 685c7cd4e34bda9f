# round-2 GPU job 17 (1 GPU): the other BASELINE configurations through bench.py (ours, stock PyTorch on the GPU, CPU reference)
for c in 2 3 4 1; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_cfg${c}_n1.log 2>&1; echo "cfg $c rc=$? $(grep '^{' gpurun_out/r2_cfg${c}_n1.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); g=d.get('gpu_reference',{}); print(round(d['value'],1), round(d['ms_per_step'],3), {k:(round(v['value'],1) if isinstance(v,dict) and 'value' in v else v) for k,v in g.items()}, round(d['roofline']['frac'],3), d['cpu_baseline']['value'])
")"
  grep -E "Error|error" gpurun_out/r2_cfg${c}_n1.log | head -3
done
