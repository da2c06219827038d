set -x
T=${TAG:-d4}
python -m pytest tests/test_motion_flowcon.py tests/test_fused_resize.py -m gpu -q 2>&1 | tail -60 > gpurun_out/${T}_pytest_new.log
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
