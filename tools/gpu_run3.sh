#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary3.txt
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -k "vae_forward or discriminator" -s > gpurun_out/m1_nets.log 2>&1; echo "nets rc=$?" >> gpurun_out/summary3.txt
timeout 2400 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -s -k "training_step or round_trip" > gpurun_out/m2_steps.log 2>&1; echo "steps rc=$?" >> gpurun_out/summary3.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary3.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary3.txt
cat gpurun_out/summary3.txt; grep -E "^\[|passed|failed" gpurun_out/m1_nets.log | tail -20; grep -E "^\[|passed|failed|Error" gpurun_out/m2_steps.log | tail -40; tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench1.log
