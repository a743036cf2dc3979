mkdir -p gpurun_out
N=${1:-8}
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload cfg5shard --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r1g_mg_bench_cfg5_$N.log 2>&1
tail -c 600 gpurun_out/r1g_mg_bench_cfg5_$N.log
