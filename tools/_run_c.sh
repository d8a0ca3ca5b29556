mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py -m gpu -x -q -s > gpurun_out/r2_c_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c_train.log
tail -40 gpurun_out/r2_c_train.log
