mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_w_pytest.log
tail -3 gpurun_out/r2_w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
