#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=...]...  -> variants/libdex_NAME.so (tuning experiments; git-ignored)
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o variants/libdex_$name.so dexterity_b200/csrc/dex_api.cu
