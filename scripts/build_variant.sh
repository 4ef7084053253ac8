#!/bin/bash
# usage: scripts/build_variant.sh TAG [-DFLAG ...]   ->  build/libphmm_TAG.so (tuning variants; never shipped)
tag=$1; shift
cd "$(dirname "$0")/../nanopore_b200/csrc" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 \
  -Xcompiler -fPIC -shared "$@" -o ../../build/libphmm_$tag.so phmm_api.cu
