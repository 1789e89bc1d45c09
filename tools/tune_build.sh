#!/bin/bash
# Builds tuning variants of the library: tools/tune_build.sh <tag> <extra nvcc flags...>
# -> materialist_b200/tuning/lib_<tag>.so (select with MB200_LIB=...)
set -e
cd "$(dirname "$0")/../materialist_b200/csrc"
tag=$1; shift
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -prec-div=false -prec-sqrt=false -Xcompiler -fPIC -ccbin /usr/bin/g++"
mkdir -p ../tuning
$NV "$@" -shared -o ../tuning/lib_${tag}.so mb200_api.cu mb200_env.cu mb200_render.cu mb200_mesh.cu mb200_lanes.cu mb200_posmlp.cu mb200_posmlp_tc.cu mb200_envutils.cu mb200_optim.cu mb200_io.cu -lz -Xptxas -v 2> ../tuning/lib_${tag}.ptxas.log
mkdir -p ../tuning; grep -E "shade_(fwd|bwd)_kernelILi1ELb(0|1)ELb0ELb0|shade_fwd_kernelILi1ELb0" -A2 ../tuning/lib_${tag}.ptxas.log | grep -E "Used|spill" | head -8
