#!/bin/bash
# usage: tools/build_variants.sh MACRO v1 v2 ...   -> build/var/lib_<MACRO>_<v>.so (kernel experiments)
M=$1; shift
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -D$M=$v \
    -o build/var/lib_${M}_$v.so ska_pst_dsp_model_b200/csrc/pfb_capi.cu 2>/dev/null &
done
wait
ls -la build/var
