# Round-end style run on a GPU box: parity tests, benches, launch list (outputs under gpurun_out/).
set -x
T=${TAG:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python bench.py --steps 10 --warmup 3 --layout nhwc --no-cpu-baseline > gpurun_out/${T}_bench_nhwc.json 2> gpurun_out/${T}_bench_nhwc.err
python bench.py --steps 10 --warmup 3 --layout nchw --no-cpu-baseline > gpurun_out/${T}_bench_nchw.json 2> gpurun_out/${T}_bench_nchw.err
python tools/prof_one.py --layout nhwc --frames 40 --iters 2 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_nhwc.csv python tools/prof_one.py --layout nhwc --frames 40 --iters 2 > gpurun_out/${T}_ncu.log 2>&1
