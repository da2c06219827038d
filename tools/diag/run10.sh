set -x
T=${TAG:-d10}
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
C2M_STRESS_DET=1 python tools/stress.py 300 13 > gpurun_out/${T}_stress_det.log 2>&1
python tools/stress.py 300 17 > gpurun_out/${T}_stress.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-other-layout"
$B > gpurun_out/${T}_bench.json 2>gpurun_out/${T}_bench.err
$B --deterministic --no-configs > gpurun_out/${T}_bench_det.json 2>/dev/null
# where does the deterministic gather spend its time
ncu --set full --import-source on --clock-control none -k regex:gather_nhwc -s 1 -c 1 -o gpurun_out/${T}_det_full python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --deterministic > gpurun_out/${T}_ncu.log 2>&1
