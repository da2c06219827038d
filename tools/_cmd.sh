timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1o_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --layout nchw --no-other-layout > gpurun_out/r1o_bench_nchw.json 2> gpurun_out/r1o_bench_nchw.err
