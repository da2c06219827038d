set -x
T=${TAG:-d6}
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
python tools/stress.py 400 7 > gpurun_out/${T}_stress.log 2>&1
C2M_STRESS_DET=1 python tools/stress.py 300 11 > gpurun_out/${T}_stress_det.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-other-layout"
$B > gpurun_out/${T}_bench.json 2>gpurun_out/${T}_bench.err
$B --deterministic --no-configs > gpurun_out/${T}_bench_det.json 2>/dev/null
C2M_WARP_FLEX=0 $B > gpurun_out/${T}_bench_noflex.json 2>/dev/null
