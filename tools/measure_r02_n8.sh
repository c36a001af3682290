# 8-GPU lines of the last build: C4 (matrix sets sharded over the ranks) and C3 (driver's launch form).
set -x
O=gpurun_out/r2n8; mkdir -p $O
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --workload c4 --steps 8 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err
tail -c 600 $O/bench_c4_n$N.err $O/bench_c3_n$N.err
