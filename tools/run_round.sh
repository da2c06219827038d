set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1_pytest.log
python bench.py --steps 10 --warmup 3 --layout nchw > gpurun_out/r1_bench_nchw.json 2> gpurun_out/r1_bench_nchw.err
python bench.py --steps 10 --warmup 3 --layout nhwc --no-cpu-baseline > gpurun_out/r1_bench_nhwc.json 2> gpurun_out/r1_bench_nhwc.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
nproc > gpurun_out/r1_nproc.txt; lscpu | head -20 >> gpurun_out/r1_nproc.txt
