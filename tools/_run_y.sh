mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_y_pytest.log
tail -3 gpurun_out/r2_y_pytest.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:decoder_tc_kernel -c 1 --csv --log-file gpurun_out/r2_y_ncu_dram256_fp16f8.csv python tools/dec_once.py 256 fp16f8 1 > /dev/null 2>&1; tail -4 gpurun_out/r2_y_ncu_dram256_fp16f8.csv | cut -d, -f13-
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -c 1 -o gpurun_out/r2_y_decoder_fp16f8 python tools/dec_once.py 48 fp16f8 1 > /dev/null 2>&1; ls -la gpurun_out/r2_y_decoder_fp16f8.ncu-rep
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_y_bench.json 2> gpurun_out/r2_y_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_y_bench.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['max_abs_vs_golden'], d['dtype'][:8])
print(d['alt_precision']); print(d['e2e_api']['value'], d['sparse']['noisy']['generate_mesh_ms'], d['sparse']['smooth']['generate_mesh_ms'], d['train']['ms_per_step'])
PY
tail -3 gpurun_out/r2_y_bench.err
