set -x
O=gpurun_out/r2z; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1
python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
tail -3 $O/smoke.txt $O/pytest.txt; head -c 600 $O/bench_c3.json
