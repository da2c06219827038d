set -x
T=${TAG:-d13}
python -m pytest tests/test_fused_resize.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-configs"
$B > gpurun_out/${T}_bench.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 12 --csv --log-file gpurun_out/${T}_l.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --fwd-only > /dev/null 2>&1
