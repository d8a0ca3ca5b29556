mkdir -p gpurun_out
for v in l2h0 "" l2h3; do
  if [ -n "$v" ]; then export S3D_LIB=tools/_bin/$v/lib.so; else unset S3D_LIB; fi
  echo "== variant ${v:-l2h1(default)}"
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:decoder_tc_kernel -c 1 python tools/dec_once.py 256 fp16f8 1 2>&1 | grep -E "dram__|lts__|done"
  timeout 200 python tools/dec_bench.py 256 fp16f8 2>&1 | tail -1
done
