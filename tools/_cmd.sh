timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2l_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --layout nchw --no-other-layout > gpurun_out/r2l_bench_nchw.json 2> gpurun_out/r2l_bench_nchw.err
