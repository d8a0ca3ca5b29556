mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "tc3 or dense_grid or border" > gpurun_out/r2_q.log 2>&1; echo "rc=$?" >> gpurun_out/r2_q.log
grep -n "max-abs\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2_q.log | cut -c1-200 | tail -24
python tools/dec_bench.py 256 fp16f8,fp16x3 > gpurun_out/r2_q_decbench.log 2>&1; cat gpurun_out/r2_q_decbench.log
