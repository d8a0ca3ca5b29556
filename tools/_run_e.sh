mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "tc3 or dense or ragged or border or batched or full_size" > gpurun_out/r2_e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_e_pytest.log
grep -n "max-abs\|passed\|failed\|rc=\|Error" gpurun_out/r2_e_pytest.log | tail -24
python tools/dec_bench.py 256 fp16x3,bf16 > gpurun_out/r2_e_decbench.log 2>&1; cat gpurun_out/r2_e_decbench.log
