# Development aid (GPU box): the pair stage with and without pair_records_kernel -- same hashes, times side by side.
O=gpurun_out/records; mkdir -p $O
python tools/pair_records_check.py > $O/check_off.txt 2>&1
ECC_PAIR_RECORDS=1 python tools/pair_records_check.py > $O/check_on.txt 2>&1
cat $O/check_off.txt $O/check_on.txt
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4_off.json 2> $O/bench_c4_off.err
ECC_PAIR_RECORDS=1 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4_on.json 2> $O/bench_c4_on.err
python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1
tail -n 3 $O/pytest.txt
python - <<'PY'
import json
for k in ("off", "on"):
    d = json.load(open(f"gpurun_out/records/bench_c4_{k}.json"))
    print(k, d["ms_per_step"], d["e2e"]["ms_per_step"], d["stages"]["pair_kernel_ms_per_launch_rank0"], d["stages"]["kernel_share_of_step"], d["stages"]["mean_unperturbed"], d["stages"]["mean_perturbed_min"])
PY
