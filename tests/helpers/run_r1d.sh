mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r1k_ops.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "iterations_match or full_solve or prepare or config4_structure" 2>&1 | tail -5) > gpurun_out/r1k_solver.log 2>&1
python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1k_bench_cfg3_invh.log 2>&1
CLRSDP_INVH=0 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1k_bench_cfg3_noinvh.log 2>&1
python bench.py --workload cfg5shard --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1k_bench_cfg5_invh.log 2>&1
cat gpurun_out/r1k_ops.log gpurun_out/r1k_solver.log
