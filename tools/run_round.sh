# Round-end style run on a GPU box: smoke, parity tests, both bench arms, then the ncu evidence for the same bench
# command -- a launch list (gpu__time_duration + DRAM bytes per launch) and one `--set full` capture of the top kernels,
# each only after the same command exited 0 without ncu.  Outputs under gpurun_out/ (copy what is to be judged
# into profiles/).
set -x
T=${TAG:-r2}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "rc=$?" >> gpurun_out/${T}_bench_ref.err
( time python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench.time; echo "rc=$?" >> gpurun_out/${T}_bench.err
python bench.py --layout nchw --no-cpu-baseline --no-pyramids --no-configs --e2e-steps 0 --torch-cuda-steps 0 > gpurun_out/${T}_bench_nchw.json 2> /dev/null
python bench.py --deterministic --no-cpu-baseline --no-pyramids --no-configs --e2e-steps 0 --torch-cuda-steps 0 --no-other-layout > gpurun_out/${T}_bench_det.json 2> /dev/null
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic --e2e-steps 0 --no-other-layout --no-pyramids --no-configs --torch-cuda-steps 0"
$BENCH > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv $BENCH > gpurun_out/${T}_ncu_launches.log 2>&1
$BENCH > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'gather_nhwc|fwd_nhwc|segbin' -s 12 -c 3 \
    -o gpurun_out/${T}_full $BENCH > gpurun_out/${T}_ncu_full.log 2>&1
python tools/prof_pyramid.py > gpurun_out/${T}_pyramid_levels.txt 2>&1
python tools/bench_loss_site.py > gpurun_out/${T}_loss_site.txt 2>&1
python tools/bench_motion_site.py > gpurun_out/${T}_motion_site.txt 2>&1
