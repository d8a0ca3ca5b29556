mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py -m gpu -x -q -s > gpurun_out/r2_l_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2_l_train.log
grep -n "vgg loss train\|train forward (\|gradient errors (\|decoder fwd+bwd\|passed\|failed\|rc=\|Error" gpurun_out/r2_l_train.log | cut -c1-900
for m in plain plain_tf32; do python tools/train_bench.py $m 2>&1 | grep world; done | tee gpurun_out/r2_l_trainbench.log
