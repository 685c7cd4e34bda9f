# round-2 GPU job 13 (8 GPUs): final scaling points with the weight gradients on side streams
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 10 --warmup 3 --profile 0"
$T --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 $B > gpurun_out/r2_b13_n8.log 2>&1; echo "n8 $(grep '^{' gpurun_out/r2_b13_n8.log | cut -c90-200)"
$T --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 $B > gpurun_out/r2_b13_n4.log 2>&1; echo "n4 $(grep '^{' gpurun_out/r2_b13_n4.log | cut -c90-200)"
$T --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 $B --wgrad-side 0 > gpurun_out/r2_b13_n8_ws0.log 2>&1; echo "n8 wgrad_side0 $(grep '^{' gpurun_out/r2_b13_n8_ws0.log | cut -c90-200)"
tail -2 gpurun_out/r2_b13_n8.log | cut -c1-200
