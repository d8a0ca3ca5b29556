mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py -m gpu -x -q -s > gpurun_out/r2_b_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2_b_train.log
tail -40 gpurun_out/r2_b_train.log
timeout 900 python -m pytest tests -m gpu -x -q -s --deselect tests/test_train.py > gpurun_out/r2_b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_b_pytest.log
tail -15 gpurun_out/r2_b_pytest.log
python tools/dec_bench.py 256 fp16x3 > gpurun_out/r2_b_decbench.log 2>&1; cat gpurun_out/r2_b_decbench.log
