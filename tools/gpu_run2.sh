#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -x -k "vae_forward or discriminator" -s > gpurun_out/m1_nets.log 2>&1; echo "nets rc=$?" >> gpurun_out/summary2.txt
timeout 2400 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -k "training_step or round_trip" > gpurun_out/m2_steps.log 2>&1; echo "steps rc=$?" >> gpurun_out/summary2.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; tail -30 gpurun_out/m1_nets.log; tail -40 gpurun_out/m2_steps.log; tail -5 gpurun_out/smoke.log
