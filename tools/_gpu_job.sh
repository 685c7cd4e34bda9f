# round-2 GPU job 11: suite, batch 8 / 64 after the coalesced weight-gradient epilogue, other configs, default bench line
python -m pytest tests -m gpu -q > gpurun_out/r2_t11.log 2>&1; echo "suite rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_t11.log | cut -c1-200 | tail -8
grep -E "^E  " gpurun_out/r2_t11.log | cut -c1-250 | head -8
grep -E "\[white" gpurun_out/r2_t11.log | cut -c1-300
B="--steps 10 --warmup 3 --no-cpu-baseline --gpu-reference 0 --profile 0"
for gb in 8 16 64; do python bench.py --global-batch $gb $B > gpurun_out/r2_b11_gb${gb}.log 2>&1; echo "gb$gb $(tail -1 gpurun_out/r2_b11_gb${gb}.log | cut -c90-200)"; done
python tools/bench_configs.py > gpurun_out/r2_other_configs.jsonl 2>gpurun_out/r2_other_configs.err; cat gpurun_out/r2_other_configs.jsonl | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b11_default.log 2>&1; tail -1 gpurun_out/r2_b11_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('default', d['value'], d['e2e']['value'], 'frac', r['frac'], 'share', r['share_of_step'], d['clocks'])
print({k: (round(v.get('tflops', v.get('tbytes_per_s', 0)), 1), round(v['ms_per_step'], 2)) for k, v in r['families'].items() if v['ms_per_step'] > 0.3})
print(d.get('gpu_reference')); print(d.get('cpu_baseline'))"
