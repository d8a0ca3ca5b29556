mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gt.py tests/test_inputs.py -m gpu -x -q -s > gpurun_out/r2_h_gt.log 2>&1; echo "rc=$?" >> gpurun_out/r2_h_gt.log
grep -n "max-abs\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2_h_gt.log | tail -20
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gt.py --deselect tests/test_inputs.py > gpurun_out/r2_h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_h_pytest.log
tail -4 gpurun_out/r2_h_pytest.log
