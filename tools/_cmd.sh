timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log
timeout 300 python bench.py --deterministic --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2a_city_det.json 2> gpurun_out/r2a_city_det.err
timeout 300 bash tools/sweep_env.sh C2M_X 0 > gpurun_out/r2a_sweep.log 2>&1
