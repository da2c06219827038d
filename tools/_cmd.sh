timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2j_pytest.log
timeout 300 bash tools/sweep_env.sh C2M_X 0 1 > gpurun_out/r2j_sweep.log 2>&1
