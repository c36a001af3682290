#!/bin/bash
# Development aid (GPU box): time the quad Radon kernel's two paths per range of angle groups
# (ECC_ITEM_AG_LO / _HI restrict the items that are computed): texture warps alone (mode 1) and window warps alone (mode 2).
# Output: one line per (range, mode) with ms/projection; the samples per range come from tools/bank_sim.bin_lines.
cd "$(dirname "$0")/.."
export N_PROJ=${N_PROJ:-16} REPS=${REPS:-2} INTERP=2
STEP=${STEP:-8}
for lo in $(seq 0 $STEP 95); do
  hi=$((lo + STEP - 1))
  for mode in 1 2; do
    echo "ag $lo-$hi mode $mode: $(ECC_ITEM_AG_LO=$lo ECC_ITEM_AG_HI=$hi ECC_HYBRID_MODE=$mode python tools/radon_variants.py 2>&1 | tail -1)"
  done
done
