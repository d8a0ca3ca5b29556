set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/r1_e_bench_256.json 2> gpurun_out/r1_e_bench_256.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_e_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_e_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -s 2 -c 1 -o gpurun_out/r1_e_decoder_bf16x3 python tools/dec_bench.py 48 bf16x3 > gpurun_out/r1_e_ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 45 -c 45 -o gpurun_out/r1_e_conv_tc python tools/enc_prof.py > gpurun_out/r1_e_ncu_enc.log 2>&1
python tools/dec_bench.py 128 bf16x3,bf16 > gpurun_out/r1_e_decbench_128.log 2>&1
ls -la gpurun_out
cat gpurun_out/r1_e_bench_256.json
