set -x
T=${TAG:-d18}
B="python bench.py --layout nchw --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-configs --no-other-layout"
$B > gpurun_out/${T}_a.json 2>/dev/null
C2M_WARP_NCHW_GF=1 $B > gpurun_out/${T}_b.json 2>/dev/null
C2M_WARP_NCHW_GF=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nchw or full_size or random_shapes" 2>&1 | tail -2 > gpurun_out/${T}_pytest.log
