# round-2 GPU job 26 (8 GPUs): the 8-GPU point with the final code
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --profile 0 > gpurun_out/r2_j26_n8.log 2>&1; echo "n8 rc=$? $(grep '^{' gpurun_out/r2_j26_n8.log | cut -c90-200)"
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' '
