#!/bin/sh
# Build libleafk.so for sm_100a in-tree (travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
OUT=../lib
mkdir -p $OUT
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
$NVCC $FLAGS -shared -o $OUT/libleafk.so leafk_api.cu k0_banks.cu k1_fp32.cu k1_tc.cu k2_pcen.cu bwd.cu prep.cu extras.cu 2> $OUT/build.log || { cat $OUT/build.log; exit 1; }
grep -E "error|warning|Used|spill" $OUT/build.log | grep -v "^$" | head -60
echo "built $OUT/libleafk.so"
