# Round-end style run on a GPU box: smoke, parity tests, benches, ncu launch list + one full capture per
# top kernel (each only after the same command exited 0 without ncu).  Outputs under gpurun_out/.
set -x
T=${TAG:-r1}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-layout --no-pyramids --torch-cuda-steps 0"
$BENCH > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv $BENCH > gpurun_out/${T}_ncu_launches.log 2>&1
$BENCH > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'gather_nhwc|fwd_nhwc|bin_kernel|overflow' -s 12 -c 4 \
    -o gpurun_out/${T}_full $BENCH > gpurun_out/${T}_ncu_full.log 2>&1
