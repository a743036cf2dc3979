mkdir -p gpurun_out
python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_cfg3_v0.log 2>&1
CLRSDP_CARRY_VAR=1 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_cfg3_v1.log 2>&1
CLRSDP_CARRY_VAR=1 python bench.py --workload cfg5shard --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_cfg5_v1.log 2>&1
python bench.py --workload cfg5shard --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_cfg5_v0.log 2>&1
(timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3) > gpurun_out/r1g_ops.log 2>&1
(CLRSDP_CARRY_VAR=1 timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3) > gpurun_out/r1g_ops_v1.log 2>&1
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/r1g_smoke.log 2>&1
cat gpurun_out/r1g_ops.log gpurun_out/r1g_ops_v1.log gpurun_out/r1g_smoke.log
