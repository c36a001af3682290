#!/bin/bash
# Development aid (GPU box): times the Radon-kernel build variants in epipolarconsistency_b200/lib/libecc_b200.<name>.so
# (make -C epipolarconsistency_b200/csrc variant NAME=.. UNIT=ecc_radon_hybrid4 EXTRA=..) at the C2/C3 size, with the
# run-time queue (INTERP=2: balances itself) and with the static split at several window-path shares (INTERP=3).
# Usage: tools/radon_ab.sh "variant names" "split per-mille values"  > log
cd "$(dirname "$0")/.."
VARIANTS=${1:-"scalar x2 x2s"}
SPLITS=${2:-"605"}
export N_PROJ=${N_PROJ:-64} REPS=${REPS:-3}
for v in $VARIANTS; do
  lib=epipolarconsistency_b200/lib/libecc_b200.$v.so
  [ "$v" = "default" ] && lib=epipolarconsistency_b200/lib/libecc_b200.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  echo "== $v queue: $(ECC_B200_LIB=$PWD/$lib INTERP=2 python tools/radon_variants.py 2>&1 | tail -1)"
  echo "== $v window-only: $(ECC_B200_LIB=$PWD/$lib ECC_HYBRID_MODE=2 INTERP=2 python tools/radon_variants.py 2>&1 | tail -1)"
  for s in $SPLITS; do
    echo "== $v static $s: $(ECC_B200_LIB=$PWD/$lib ECC_HYBRID4_SPLIT=$s INTERP=3 python tools/radon_variants.py 2>&1 | tail -1)"
  done
done
