mkdir -p gpurun_out
for v in base/libbase.so ct_dual_nosplit/lib.so ct_single_split/lib.so ""; do
  if [ -n "$v" ]; then export S3D_LIB=tools/_bin/$v; else unset S3D_LIB; fi
  echo "== variant ${v:-default(dual+split)}"
  timeout 200 python tools/enc_check.py k12_s256_g128_g256 2>&1 | tail -7
done
unset S3D_LIB
timeout 200 python tools/enc_check.py k12_s128_g128 2>&1 | tail -7
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encoder or planes or gt or vgg" 2>&1 | tail -3
