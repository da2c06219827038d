set -x
T=${TAG:-d17}
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv"
$NCU --log-file gpurun_out/${T}_l_kitti_det.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --workload kitti_256x832_c64_oob --deterministic > /dev/null 2>&1
$NCU --log-file gpurun_out/${T}_l_det.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --deterministic > /dev/null 2>&1
