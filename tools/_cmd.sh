timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1j_pytest.log
timeout 300 bash tools/sweep_env.sh C2M_WARP_FUSED_BIN 0 1 > gpurun_out/r1j_sweep.log 2>&1
