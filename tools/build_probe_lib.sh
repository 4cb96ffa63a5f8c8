#!/bin/bash
# Development build of the library with extra -D flags into tools/probe_lib/ (git-ignored; travels with gpurun).
#   tools/build_probe_lib.sh NAME -DCM_DEV_PROBES ...   ->  tools/probe_lib/NAME.so   (use with CM_LIBPATH)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=tools/probe_lib/$name.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr --extended-lambda \
  -Xcompiler -fPIC "$@" -shared -o $out cellmapper_b200/csrc/{cabi,knn_exact,knn_mma,graph_kernel,transfer,jaccard}.cu -lcudart
echo $out
