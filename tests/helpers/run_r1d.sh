mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r1h_ops.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "iterations_match or full_solve or prepare" 2>&1 | tail -5) > gpurun_out/r1h_solver.log 2>&1
python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_cfg3_split.log 2>&1
CLRSDP_BN_SPLIT=0 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_cfg3_nosplit.log 2>&1
cat gpurun_out/r1h_ops.log gpurun_out/r1h_solver.log
