#!/bin/bash
# Development aid (GPU box): texture warps alone (mode 1), per range of angle groups and per tiling of the texture sub-tile
# (ECC_HYBRID4_TEXMAP); all tilings compute the same bins with the same arithmetic (same checksum).
cd "$(dirname "$0")/.."
export N_PROJ=${N_PROJ:-16} REPS=${REPS:-2} INTERP=2 ECC_HYBRID_MODE=1
for lo in 0 24 44 64; do
  hi=$((lo + 7))
  for map in 0 1 2 3 4 5; do
    echo "ag $lo-$hi texmap $map: $(ECC_ITEM_AG_LO=$lo ECC_ITEM_AG_HI=$hi ECC_HYBRID4_TEXMAP=$map python tools/radon_variants.py 2>&1 | tail -1 | sed 's/.*kernel only/kernel only/')"
  done
done
