mkdir -p gpurun_out
for v in gb4 gb8 "" gb32; do
  if [ -n "$v" ]; then export S3D_LIB=tools/_bin/$v/lib.so; else unset S3D_LIB; fi
  echo "== variant ${v:-gb16(default)}"
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:decoder_tc_kernel -c 1 python tools/dec_once.py 256 fp16f8 1 2>&1 | grep -E "dram__|lts__|done"
done
