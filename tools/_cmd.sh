python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1g_pytest.log
echo "occ3" > gpurun_out/r1g_sweep.log
bash tools/sweep_env.sh C2M_X 0 >> gpurun_out/r1g_sweep.log 2>&1
sed -i "s/__launch_bounds__(256, [0-9]) gather_nhwc_kernel/__launch_bounds__(256, 4) gather_nhwc_kernel/" c2m_b200/csrc/warp_bwd_gather.cu
python -m c2m_b200._build >> gpurun_out/r1g_sweep.log 2>&1
echo "occ4" >> gpurun_out/r1g_sweep.log
bash tools/sweep_env.sh C2M_X 0 >> gpurun_out/r1g_sweep.log 2>&1
