mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_d_pytest.log
grep -n "passed\|failed\|rc=\|Error" gpurun_out/r2_d_pytest.log | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err; tail -c 6000 gpurun_out/r2_d_bench.json; tail -5 gpurun_out/r2_d_bench.err
