mkdir -p gpurun_out
for i in 1 2; do
echo "== base"; S3D_LIB=tools/_bin/base/libbase.so timeout 200 python tools/dec_bench.py 256 fp16f8,bf16 2>&1 | tail -4
echo "== new"; timeout 200 python tools/dec_bench.py 256 fp16f8,bf16 2>&1 | tail -4
done
