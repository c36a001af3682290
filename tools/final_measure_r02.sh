set -x
O=gpurun_out/r5a; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1
python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
for w in c1 c2 c4 c5; do python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; done
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python tools/direct_bench.py 100 > $O/direct_bench.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:radon_hybrid4_kernel -s 2 -c 1 -o $O/radon_hybrid4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_radon.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^void .*::pairs_kernel" -s 3 -c 1 -o $O/pairs_c5 python tools/c5_pair_probe.py > $O/ncu_pairs_c5.log 2>&1
ls -la $O
