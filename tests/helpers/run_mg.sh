mkdir -p gpurun_out
N=${1:-2}   # usage: gpurun --gpus N -- "bash tests/helpers/run_mg.sh N": 2-GPU parity check + weak-scaling benches
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -12) > gpurun_out/r1f_mg_check_$N.log 2>&1
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r1f_mg_bench_cfg3_$N.log 2>&1
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload cfg5shard --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r1f_mg_bench_cfg5_$N.log 2>&1
tail -2 gpurun_out/r1f_mg_check_$N.log
