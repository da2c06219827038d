set -x
T=${TAG:-d2}
python tools/probe_affine.py > gpurun_out/${T}_probe_affine.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_resolution" 2>&1 | tail -15 > gpurun_out/${T}_pytest_fullres.log
export LD_PRELOAD=$PWD/tools/diag/libterm_trace.so
python tools/stress.py 600 1 > gpurun_out/${T}_stress.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_stress.log
python tools/bench_loss_site.py > gpurun_out/${T}_loss.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_loss.log
