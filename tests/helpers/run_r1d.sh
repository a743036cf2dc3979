mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r1i_tests.log 2>&1
python bench.py > gpurun_out/r1i_bench.log 2>&1
CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_plain.log 2>&1 && \
CLRSDP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 660 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_ncu.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mma_planes -s 90 -c 6 -o gpurun_out/r1i_mma -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_ncu2.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"carry_kernel|slice_rows|panel_factor" -s 200 -c 12 -o gpurun_out/r1i_others -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_ncu3.log 2>&1
tail -3 gpurun_out/r1i_tests.log
