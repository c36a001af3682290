set -x
O=gpurun_out/r2p; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1
python bench.py --workload c4 --steps 8 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
ncu --set full --clock-control none --import-source on -k regex:pairs_kernel -s 2 -c 1 -o $O/pairs_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_pairs_c4.log 2>&1
tail -n 2 $O/pytest.txt
