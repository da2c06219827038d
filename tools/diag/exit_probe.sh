# Development aid: does a process that used the library leave through a normal interpreter exit?
set -x
T=${TAG:-d1}
export LD_PRELOAD=$PWD/tools/diag/libterm_trace.so
python tools/stress.py 40 1 > gpurun_out/${T}_stress.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_stress.log
python bench.py --steps 5 --warmup 3 --no-traffic > gpurun_out/${T}_bench_small.json 2> gpurun_out/${T}_bench_small.err; echo "rc=$?" >> gpurun_out/${T}_bench_small.err
unset LD_PRELOAD
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "rc=$?" >> gpurun_out/${T}_bench_ref.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
