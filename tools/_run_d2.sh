mkdir -p gpurun_out
for v in base/libbase.so ""; do
  if [ -n "$v" ]; then export S3D_LIB=tools/_bin/$v; n=base; else unset S3D_LIB; n=new; fi
  echo "== variant $n"
  ENC_PROF_TIME=1 timeout 200 python tools/enc_prof.py 256 2>&1 | tail -1
  timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_d2_enc_launches_$n.csv python tools/enc_prof.py 256 > /dev/null 2>&1
  wc -l gpurun_out/r2_d2_enc_launches_$n.csv
done
