set -x
T=${TAG:-d5}
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python tools/stress.py 300 7 > gpurun_out/${T}_stress.log 2>&1
C2M_STRESS_DET=1 python tools/stress.py 200 11 > gpurun_out/${T}_stress_det.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-configs --no-other-layout"
$B --deterministic > gpurun_out/${T}_bench_det.json 2>/dev/null
$B > gpurun_out/${T}_bench.json 2>/dev/null
