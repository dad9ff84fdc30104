#!/bin/bash
# Build tuning variants of libvrestir.so (different -D knobs for vr_wavefront.cu) into build_variants/; select one at
# run time with VRESTIR_LIB=<path> (developer override read by _capi.py).
# usage: tools/build_variants.sh name1:"-DA=1 -DB=2" name2:"..." ...
set -e
cd "$(dirname "$0")/../volumetricrestirrelease_b200/csrc"
make -j8 >/dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( nvcc $FLAGS $defs -c vr_wavefront.cu -o build/vr_wavefront_$name.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_variants/libvrestir_$name.so build/vr_kernels.o build/vr_wavefront_$name.o build/vr_pass.o build/vr_post.o build/vr_mipbuild.o build/vr_scene.o -Xlinker -lpthread && echo built $name ) &
done
wait
