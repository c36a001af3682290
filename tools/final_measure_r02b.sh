# Final single-GPU measurements of round 2 after the pair-order change (everything else as tools/final_measure_r02.sh).
set -x
O=gpurun_out/r2f; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1
python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
for w in c1 c2 c4 c5; do python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; done
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches_c4.log 2>&1
ls -la $O
