timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1t_pytest.log
timeout 300 bash tools/sweep_env.sh C2M_X 0 > gpurun_out/r1t_sweep.log 2>&1
