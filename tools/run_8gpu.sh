# 8-GPU lines of the headline benchmark (what the driver's scaling run does at N = 8); outputs under gpurun_out/
T=${TAG:-r2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err; echo "rc=$?" >> gpurun_out/${T}_bench_8gpu.err
