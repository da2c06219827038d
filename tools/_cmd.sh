timeout 300 bash tools/sweep_env.sh C2M_X 0 > gpurun_out/r2b_sweep.log 2>&1
