#!/bin/bash
# Parallel build of compile-time variants of liboctb200.so (development A/B only): tools/variant_build_parallel.sh "tag:DEFS" ...
# Every object of every variant is one nvcc job, JOBS at a time; then one link per variant -> octproz_b200/variants/liboctb200_<tag>.so
set -e
cd "$(dirname "$0")/../octproz_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v"
OBJS="octb200 k_aux k_generic k_fused_r1 k_fused_r2 k_fused_p12 k_fused_c8c32 k_fused_cplx k_fused_r1_conv k_fused_r2_conv k_fused_p12_conv"
[ -n "$ONLY_OBJS" ] && OBJS="$ONLY_OBJS"
JOBS=${JOBS:-8}
list=$(mktemp)
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  mkdir -p ../variants/obj_$tag
  for f in $OBJS; do
    src=${f%_conv}; def=""; [ "$src" != "$f" ] && def="-DOCT_FUSED_CONV=1"
    echo "$NVCC $FLAGS $defs $def -c $src.cu -o ../variants/obj_$tag/$f.o 2> ../variants/obj_$tag/$f.ptxas.log || (echo FAILED $tag $f; cat ../variants/obj_$tag/$f.ptxas.log | head -20)" >> $list
  done
done
# heavy translation units first
sort -t/ -k4 -r $list | xargs -P $JOBS -I{} bash -c "{}"
for spec in "$@"; do
  tag=${spec%%:*}
  ALL="octb200 k_aux k_generic k_fused_r1 k_fused_r2 k_fused_p12 k_fused_c8c32 k_fused_cplx k_fused_r1_conv k_fused_r2_conv k_fused_p12_conv"
  objs=""; for f in $ALL; do if [ -f ../variants/obj_$tag/$f.o ]; then objs="$objs ../variants/obj_$tag/$f.o"; else objs="$objs $f.o"; fi; done
  $NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/liboctb200_$tag.so $objs -ldl
  echo "linked $tag: $(grep -h spill ../variants/obj_$tag/k_fused_r1.ptxas.log | sort | uniq -c | tr '\n' ';')"
done
rm -f $list
