set -x
T=${TAG:-d14}
ncu --set full --import-source on --clock-control none -k regex:gather_nhwc -s 1 -c 1 -o gpurun_out/${T}_det_full python tools/prof_one.py --layout nhwc --frames 40 --iters 2 --deterministic > gpurun_out/${T}_ncu.log 2>&1
