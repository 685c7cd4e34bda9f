# round-2 GPU job 25 (2 GPUs): data-parallel parity and the 2-GPU point with the final code
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
DP_WIRE=fp32 timeout 300 $T 29541 tools/dp_check.py > gpurun_out/r2_j25_dp2_fp32.log 2>&1; echo "dp fp32 rc=$? $(grep '^{' gpurun_out/r2_j25_dp2_fp32.log | cut -c1-260)"
DP_WIRE=bf16 timeout 300 $T 29542 tools/dp_check.py > gpurun_out/r2_j25_dp2_bf16.log 2>&1; echo "dp bf16 rc=$? $(grep '^{' gpurun_out/r2_j25_dp2_bf16.log | cut -c1-260)"
timeout 300 $T 29543 bench.py --gpus 2 --steps 10 --warmup 3 --profile 0 > gpurun_out/r2_j25_n2.log 2>&1; echo "n2 rc=$? $(grep '^{' gpurun_out/r2_j25_n2.log | cut -c90-200)"
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
