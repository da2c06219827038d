set -x
T=${TAG:-d9}
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or deterministic or out_of_bounds or channels_last" 2>&1 | tail -4 > gpurun_out/${T}_pytest.log
python tools/stress.py 300 7 > gpurun_out/${T}_stress.log 2>&1
C2M_STRESS_DET=1 python tools/stress.py 200 11 > gpurun_out/${T}_stress_det.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-other-layout"
$B > gpurun_out/${T}_bench.json 2>gpurun_out/${T}_bench.err
C2M_WARP_FLEX=0 $B --no-configs > gpurun_out/${T}_bench_noflex.json 2>/dev/null
$B --no-configs > gpurun_out/${T}_bench2.json 2>/dev/null
