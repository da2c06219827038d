set -x
T=${TAG:-d8}
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
python tools/stress.py 400 7 > gpurun_out/${T}_stress.log 2>&1
C2M_STRESS_DET=1 python tools/stress.py 300 11 > gpurun_out/${T}_stress_det.log 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-other-layout"
$B > gpurun_out/${T}_bench.json 2>gpurun_out/${T}_bench.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 80 --csv"
$NCU --log-file gpurun_out/${T}_l_kitti.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob > /dev/null 2>&1
$NCU --log-file gpurun_out/${T}_l_kitti_det.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob --deterministic > /dev/null 2>&1
