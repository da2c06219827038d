set -x
T=${TAG:-d12}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; echo "rc=$?" >> gpurun_out/${T}_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref_2gpu.json 2> gpurun_out/${T}_ref_2gpu.err; echo "rc=$?" >> gpurun_out/${T}_ref_2gpu.err
