python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_generator.py > gpurun_out/r2n_gen_g2.json 2> gpurun_out/r2n_gen_g2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_g2.json 2> gpurun_out/r2n_bench_g2.err
