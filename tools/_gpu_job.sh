# round-2 GPU job 18 (4 GPUs): BASELINE configuration 4 (Cycle-VAE unpaired, global batch 32) data-parallel on 2 and 4 GPUs
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 2 --master-port 29551 bench.py --config 4 --gpus 2 --steps 10 --warmup 3 --profile 0 > gpurun_out/r2_cfg4_n2.log 2>&1; echo "cfg4 n2 rc=$? $(grep '^{' gpurun_out/r2_cfg4_n2.log | cut -c80-230)"
timeout 300 $T --nproc-per-node 4 --master-port 29552 bench.py --config 4 --gpus 4 --steps 10 --warmup 3 --profile 0 > gpurun_out/r2_cfg4_n4.log 2>&1; echo "cfg4 n4 rc=$? $(grep '^{' gpurun_out/r2_cfg4_n4.log | cut -c80-230)"
grep -E "Error|error" gpurun_out/r2_cfg4_n2.log gpurun_out/r2_cfg4_n4.log | head -5
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
