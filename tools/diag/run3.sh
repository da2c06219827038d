set -x
T=${TAG:-d3}
python -m pytest tests/test_fused_resize.py -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest_resize.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python bench.py --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
