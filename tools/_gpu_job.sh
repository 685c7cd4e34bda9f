# round-2 GPU job 9 (8 GPUs): the headline scaling points
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 10 --warmup 3 --profile 0"
$T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 $B > gpurun_out/r2_b9_n8.log 2>&1; echo "n8 $(grep '^{' gpurun_out/r2_b9_n8.log | cut -c90-200)"
$T --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 $B > gpurun_out/r2_b9_n4.log 2>&1; echo "n4 $(grep '^{' gpurun_out/r2_b9_n4.log | cut -c90-200)"
$T --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 $B --overlap 0 > gpurun_out/r2_b9_n8_ov0.log 2>&1; echo "n8 overlap0 $(grep '^{' gpurun_out/r2_b9_n8_ov0.log | cut -c90-200)"
$T --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 $B --wire fp32 > gpurun_out/r2_b9_n8_fp32.log 2>&1; echo "n8 wire fp32 $(grep '^{' gpurun_out/r2_b9_n8_fp32.log | cut -c90-200)"
tail -2 gpurun_out/r2_b9_n8.log | cut -c1-300
