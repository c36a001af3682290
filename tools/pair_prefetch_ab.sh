# Development aid (GPU box): the pair stage of the build variants libecc_b200.<name>.so (make variant UNIT=ecc_pairs ...) at the
# C3 and C4 shapes -- hashes of all pair values and times.  Usage: tools/pair_prefetch_ab.sh "names" > log
cd "$(dirname "$0")/.."
export BINS=${BINS:-768} SETS=${SETS:-16}
for v in default $1; do
  lib=epipolarconsistency_b200/lib/libecc_b200.$v.so
  [ "$v" = "default" ] && lib=epipolarconsistency_b200/lib/libecc_b200.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  echo "== $v"
  ECC_B200_LIB=$PWD/$lib python tools/pair_records_check.py 2>&1 | tail -4
done
