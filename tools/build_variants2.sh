#!/bin/bash
# usage: tools/build_variants2.sh name "-DA=1 -DB=2" ...   -> build/var/lib_<name>.so
while [ $# -gt 1 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $2 \
    -o build/var/lib_$1.so ska_pst_dsp_model_b200/csrc/pfb_capi.cu 2>/dev/null &
  shift 2
done
wait
ls build/var
