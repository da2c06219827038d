# BASELINE configs[3] and [4] at 8 GPUs of one box (one process per GPU): generator training step under DDP,
# full-resolution C=256 shards.  Outputs under gpurun_out/.
T=${TAG:-r6}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/bench_generator.py --warp both \
    > gpurun_out/${T}_gen_ddp8.json 2> gpurun_out/${T}_gen_ddp8.err
timeout 400 $TR --master-port 29512 bench.py --gpus 8 --workload fullres_1024x2048_c256 --steps 5 --warmup 3 \
    --no-cpu-baseline --e2e-steps 0 --no-other-layout --no-pyramids --torch-cuda-steps 0 \
    > gpurun_out/${T}_fullres_g8.json 2> gpurun_out/${T}_fullres_g8.err
