#!/bin/bash
# Round-2 starter: build the compile-time variants of the fused kernel that were prepared (and counted statically) but not yet timed,
# then time each against the default build on ONE box with tools/micro_ab.py (device time per volume + output hash: a variant that
# is faster must also be bit-identical).  Step 1 runs here (no GPU, ~3 min per variant), step 2 on the box:
#   bash tools/variant_sweep.sh build
#   gpurun --timeout 300 -- 'bash tools/variant_sweep.sh run'
# Round-2 result (profiles/r02a_variant_sweep.txt): TW4 + x32 reads became the N = 1024 default (0.2114 -> 0.2042 ms), TW4 + the
# no-shift conversion the N = 2048 default; what is listed below is each remaining switch on its own against the new default.
set -e
cd "$(dirname "$0")/.."
declare -A V=(
  [r1old]="-DOCT_R1_TW4=0 -DOCT_R1_XCHG_X=8 -DOCT_R2_TW4=0 -DOCT_R2_NOSHIFT=0"
  [r1i2f]="-DOCT_CVT_I2F=1"
  [r1epi]="-DOCT_EPI_PAIR=1"
  [r1w20]="-DOCT_R1_THREADS=640"
  [r2x16]="-DOCT_R2_XCHG_X=16"
  [r2tw2]="-DOCT_R2_TW4=0"
  [r2egvar]="-DOCT_R2_EGVAR=1"
)
case "$1" in
  build)
    specs=(); for t in "${!V[@]}"; do specs+=("$t:${V[$t]}"); done
    ONLY_OBJS="k_fused_r1 k_fused_r2 k_fused_r1_conv k_fused_r2_conv k_fused_p12 k_fused_p12_conv" tools/variant_build_parallel.sh "${specs[@]}"
    rm -rf octproz_b200/variants/obj_* ;;
  run)
    python tools/micro_ab.py default 1024; python tools/micro_ab.py default 2048
    for t in "${!V[@]}"; do
      for n in 1024 2048; do OCTB200_LIB=$PWD/octproz_b200/variants/liboctb200_$t.so python tools/micro_ab.py $t $n || echo "variant $t N=$n failed"; done
    done
    python tools/micro_ab.py default_again 1024 ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
