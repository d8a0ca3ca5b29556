mkdir -p gpurun_out
# (1) launch list of the bench step (inference): kernel shares
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-train --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2_launches.csv | cut -c1-300
# (2) full capture of the decoder kernel (48^3 grid, one launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -c 1 -f -o gpurun_out/r2_decoder_fp16x3 python tools/dec_once.py 48 fp16x3 1 > gpurun_out/r2_ncu_dec.log 2>&1; tail -2 gpurun_out/r2_ncu_dec.log
# (3) launch list of one training step (second step of two)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_once.py 2 > gpurun_out/r2_train_under_ncu.log 2>&1
wc -l gpurun_out/r2_train_launches.csv
# (4) launch list of the sparse (MISE) + marching cubes pipeline
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_mise|k_mc_|decoder_tc" -c 2000 --csv --log-file gpurun_out/r2_mesh_launches.csv python tools/mesh_bench.py > gpurun_out/r2_mesh_under_ncu.log 2>&1
wc -l gpurun_out/r2_mesh_launches.csv
# (5) the numbers themselves, not under a profiler
python tools/mesh_bench.py > gpurun_out/r2_mesh_bench.log 2>&1; cat gpurun_out/r2_mesh_bench.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_j_bench.json 2> gpurun_out/r2_j_bench.err; cut -c1-400 gpurun_out/r2_j_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_j_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2_j_bench_ref.json
