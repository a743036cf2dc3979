#!/bin/bash
# One gpurun call that produces everything a round records (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tests/helpers/gpu_round_check.sh r2a'
# -> gpurun_out/<tag>_tests.log (pytest -m gpu), <tag>_bench.log (default bench line with cpu_baseline),
#    <tag>_launches.csv (ncu launch list), <tag>_mma.ncu-rep and <tag>_others.ncu-rep (ncu --set full);
# then, back in the container:  python profiles/make_summaries.py <tag>
TAG=${1:-check}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py > gpurun_out/${TAG}_bench.log 2>&1
CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 && \
CLRSDP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 660 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mma_planes -s 90 -c 6 -o gpurun_out/${TAG}_mma -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"carry_kernel|slice_rows|panel_factor" -s 200 -c 12 -o gpurun_out/${TAG}_others -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu3.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
