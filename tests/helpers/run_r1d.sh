mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r1f_tests.log 2>&1
python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_cfg3.log 2>&1
python bench.py --workload cfg5shard --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_cfg5.log 2>&1
tail -3 gpurun_out/r1f_tests.log
