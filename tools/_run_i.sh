mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "ragged_empty or batched or tc3" > gpurun_out/r2_i_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2_i_k.log
grep -n "max-abs\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2_i_k.log | tail -14
python tools/dec_bench.py 256 fp16x3 > gpurun_out/r2_i_decbench.log 2>&1; tail -1 gpurun_out/r2_i_decbench.log
