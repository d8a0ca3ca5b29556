#!/bin/bash
# Build an experiment variant of the library: tools/build_variant.sh NAME "-DMACRO=1 ..." [file.cu ...]  -> tools/_bin/NAME/lib.so
# (only the named sources are recompiled with the extra flags; the other objects come from slice3d_b200/_lib)
set -e
name=$1; flags=$2; shift 2
srcs=${@:-decoder_tc.cu}
d=tools/_bin/$name; mkdir -p $d
objs=""
for o in slice3d_b200/_lib/*.o; do
  b=$(basename $o .o)
  if [[ " $srcs " == *" $b.cu "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c slice3d_b200/csrc/$b.cu -o $d/$b.o
    objs="$objs $d/$b.o"
  else
    objs="$objs $o"
  fi
done
nvcc -shared -o $d/lib.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static
echo $d/lib.so
