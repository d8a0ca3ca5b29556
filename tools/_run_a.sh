mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_a_pytest.log
tail -5 gpurun_out/r2_a_pytest.log
python tools/dec_bench.py 256 fp16x3,bf16x3,bf16 > gpurun_out/r2_a_decbench.log 2>&1
cat gpurun_out/r2_a_decbench.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:decoder_tc_kernel -c 1 --csv --log-file gpurun_out/r2_a_ncu_dram256.csv python tools/dec_once.py 256 fp16x3 1 > gpurun_out/r2_a_ncu_dram256.log 2>&1
cat gpurun_out/r2_a_ncu_dram256.csv | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err; tail -c 3000 gpurun_out/r2_a_bench.json
