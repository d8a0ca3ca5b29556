mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_mcubes.py -m gpu -x -q -s -k "fp16_plus_fp8 or marching" > gpurun_out/r2_p.log 2>&1; echo "rc=$?" >> gpurun_out/r2_p.log
grep -n "fp16 + 2\|passed\|failed\|rc=\|Error\|error\|generate_mesh" gpurun_out/r2_p.log | cut -c1-300
