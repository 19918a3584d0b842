#!/bin/sh
# Build libleafk.so for sm_100a in-tree (travels to the GPU box with the snapshot).  The translation units compile in
# parallel (the tcgen05 kernel has 17 instantiations, split over four of them), then one link.
set -e
cd "$(dirname "$0")"
OUT=../lib
OBJ=$OUT/obj
mkdir -p $OUT $OBJ
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
SRCS="leafk_api k0_banks k1_fp32 k1_tc k1_tc_inst_a k1_tc_inst_b k1_tc_inst_c k1_tc_inst_t k2_pcen bwd prep extras"
for s in $SRCS; do
  ( $NVCC $FLAGS -c $s.cu -o $OBJ/$s.o > $OBJ/$s.log 2>&1 || echo FAILED > $OBJ/$s.failed ) &
done
wait
: > $OUT/build.log
for s in $SRCS; do cat $OBJ/$s.log >> $OUT/build.log; done
if ls $OBJ/*.failed > /dev/null 2>&1; then rm -f $OBJ/*.failed; cat $OUT/build.log; exit 1; fi
OBJS=""
for s in $SRCS; do OBJS="$OBJS $OBJ/$s.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libleafk.so $OBJS
grep -E "error|warning|Used|spill" $OUT/build.log | grep -v "^$" | head -60
echo "built $OUT/libleafk.so"
