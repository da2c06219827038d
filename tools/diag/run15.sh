set -x
T=${TAG:-d15}
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "deterministic or full_size or channels_last or sliced" 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
C2M_STRESS_DET=1 python tools/stress.py 400 23 > gpurun_out/${T}_stress_det.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-other-layout"
$B --deterministic > gpurun_out/${T}_bench_det.json 2>/dev/null
