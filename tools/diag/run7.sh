set -x
T=${TAG:-d7}
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 80 --csv"
$NCU --log-file gpurun_out/${T}_l_det.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --deterministic > /dev/null 2>&1
$NCU --log-file gpurun_out/${T}_l_kitti.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob > /dev/null 2>&1
$NCU --log-file gpurun_out/${T}_l_kitti_det.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob --deterministic > /dev/null 2>&1
C2M_WARP_FLEX=0 $NCU --log-file gpurun_out/${T}_l_kitti_noflex.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob > /dev/null 2>&1
