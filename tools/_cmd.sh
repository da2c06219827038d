timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1n_pytest.log
timeout 300 bash tools/sweep_env.sh C2M_WARP_BWD_LISTS 0 > gpurun_out/r1n_sweep.log 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-layout"
$BENCH > gpurun_out/r1n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 10 -c 8 --csv \
    --log-file gpurun_out/r1n_launches.csv $BENCH > gpurun_out/r1n_ncu_launches.log 2>&1
