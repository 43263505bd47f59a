#!/bin/bash
# Round-2 starter: build the compile-time variants of the fused kernel that were prepared (and counted statically) but not yet timed,
# then time each against the default build on ONE box with tools/micro_ab.py (device time per volume + output hash: a variant that
# is faster must also be bit-identical).  Step 1 runs here (no GPU, ~3 min per variant), step 2 on the box:
#   bash tools/variant_sweep.sh build
#   gpurun --timeout 300 -- 'bash tools/variant_sweep.sh run'
# Static instruction counts of the N = 1024 benchmark kernel (common part of the loop, default = 1085): x32 1022, tw4 1041,
# tw4+x32 1001, i2f 1052.
set -e
cd "$(dirname "$0")/.."
declare -A V=(
  [x32]="-DOCT_XCHG_X=32"
  [x16]="-DOCT_XCHG_X=16"
  [tw4]="-DOCT_TW4=1"
  [tw4x32]="-DOCT_TW4=1 -DOCT_XCHG_X=32"
  [tw4x32w20]="-DOCT_TW4=1 -DOCT_XCHG_X=32 -DOCT_R1_THREADS=640"
  [i2f]="-DOCT_CVT_I2F=1"
  [epipair]="-DOCT_EPI_PAIR=1"
  [all]="-DOCT_TW4=1 -DOCT_XCHG_X=32 -DOCT_CVT_I2F=1 -DOCT_EPI_PAIR=1"
  [r2noshift]="-DOCT_R2_NOSHIFT=1"
  [r2egvar]="-DOCT_R2_EGVAR=1"
)
case "$1" in
  build)
    for t in "${!V[@]}"; do echo "== $t: ${V[$t]}"; make -s -C octproz_b200/csrc variant TAG=$t DEFS="${V[$t]}" 2>&1 | grep -v deprecated || true
      grep -h "spill" octproz_b200/variants/obj_$t/k_fused_r1.ptxas.log | sort | uniq -c | head -3; done ;;
  run)
    python tools/micro_ab.py default 1024; python tools/micro_ab.py default 2048
    for t in "${!V[@]}"; do
      for n in 1024 2048; do OCTB200_LIB=$PWD/octproz_b200/variants/liboctb200_$t.so python tools/micro_ab.py $t $n || echo "variant $t N=$n failed"; done
    done
    python tools/micro_ab.py default_again 1024 ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
