mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r1d_tests.log 2>&1
python bench.py > gpurun_out/r1d_bench.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1d_bench_ref.log 2>&1
CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_plain.log 2>&1 && \
CLRSDP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 640 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mma_planes -s 90 -c 6 -o gpurun_out/r1d_mma -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu2.log 2>&1
tail -3 gpurun_out/r1d_tests.log
