# Development aid (GPU box): bench.py --workload c4 (and c3's pair stage) for build variants libecc_b200.<name>.so.
cd "$(dirname "$0")/.."
O=gpurun_out/pf; mkdir -p $O
for v in default $1; do
  lib=epipolarconsistency_b200/lib/libecc_b200.$v.so
  [ "$v" = "default" ] && lib=epipolarconsistency_b200/lib/libecc_b200.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  ECC_B200_LIB=$PWD/$lib python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_c4_$v.json 2> $O/bench_c4_$v.err
  python - $v <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/pf/bench_c4_{sys.argv[1]}.json"))
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "pair kernel ms/launch", round(d["stages"]["pair_kernel_ms_per_launch_rank0"], 3),
      "means", d["stages"]["mean_unperturbed"], d["stages"]["mean_perturbed_min"], d["stages"]["mean_perturbed_max"])
PY
done
