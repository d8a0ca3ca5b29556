mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_f2_bench.json 2> gpurun_out/r2_f2_bench.err; tail -2 gpurun_out/r2_f2_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_f2_bench_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/r2_f2_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_f2_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > /dev/null 2>&1; wc -l gpurun_out/r2_f2_launches.csv
timeout 300 compute-sanitizer --tool memcheck python tools/enc_prof.py 64 > gpurun_out/r2_f2_enc_memcheck.log 2>&1; tail -3 gpurun_out/r2_f2_enc_memcheck.log
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_f2_bench.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['max_abs_vs_golden'], d['dtype'][:8], d['variant']['precision_selection'])
print(d['e2e_api']['value'], d['sparse']['noisy']['generate_mesh_ms'], d['sparse']['smooth']['generate_mesh_ms'], d['train']['ms_per_step'], d['configs1_128']['value'], d['clocks'])
PY
