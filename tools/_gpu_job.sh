# round-2 GPU job 10 (8 GPUs): bucket tails on 3 side streams vs 1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 10 --warmup 3 --profile 0"
$T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 $B --side-streams 1 > gpurun_out/r2_b10_n8_s1.log 2>&1; echo "n8 side1 $(grep '^{' gpurun_out/r2_b10_n8_s1.log | cut -c90-200)"
$T --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 $B > gpurun_out/r2_b10_n8_s3.log 2>&1; echo "n8 side3 $(grep '^{' gpurun_out/r2_b10_n8_s3.log | cut -c90-200)"
$T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 $B --side-streams 2 > gpurun_out/r2_b10_n8_s2.log 2>&1; echo "n8 side2 $(grep '^{' gpurun_out/r2_b10_n8_s2.log | cut -c90-200)"
tail -2 gpurun_out/r2_b10_n8_s3.log | cut -c1-200
