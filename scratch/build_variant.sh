#!/bin/bash
# scratch/build_variant.sh <name> [-DFOO=1 ...]  ->  scratch/libs/<name>.so  (A/B builds of the library; selected with NRMC_RT_LIB)
name=$1; shift
unset CC CXX
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --ftz=false --prec-div=true \
  --prec-sqrt=true --fmad=true "$@" -o scratch/libs/$name.so nuradiomc_b200/csrc/nrmc_rt.cu
