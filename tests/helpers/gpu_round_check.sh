#!/bin/bash
# One gpurun call that produces everything a round records (run from the repo root on the GPU box):
#   gpurun --timeout 1800 -- 'bash tests/helpers/gpu_round_check.sh r3a'
# -> gpurun_out/<tag>_tests.log (pytest -m gpu), <tag>_bench.log (default bench line: cpu_baseline, second roofline),
#    <tag>_launches.csv (ncu launch list), <tag>_gemm256 / _gemm512 / _cuda .ncu-rep.csv (ncu --set full, exported to CSV
#    on the box: the reports themselves exceed the 64 MiB a call may copy back);
# then, back in the container:  python profiles/make_summaries.py <tag> ; python profiles/make_sass_excerpt.py <tag>
# (about 25 GPU-minutes: 9 for the tests, 3 for the bench, 11 for the captures).
TAG=${1:-check}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log 2>&1
python bench.py > gpurun_out/${TAG}_bench.log 2>&1
CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-roofline > gpurun_out/${TAG}_plain.log 2>&1 && \
CLRSDP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 640 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-roofline > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --cache-control none -k regex:mma_planes -c 20 -f -o gpurun_out/${TAG}_gemm256 python tests/gpu_micro_gemm.py 256 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --cache-control none -k regex:mma_planes -c 16 -f -o gpurun_out/${TAG}_gemm512 python tests/gpu_micro_gemm.py 512 cfg5 > gpurun_out/${TAG}_ncu3.log 2>&1
CLRSDP_GRAPH=0 ncu --set full --clock-control none -k regex:"slice_rows|panel_factor|gemv|small_gemm|lambda_min" -s 300 -c 24 -f -o gpurun_out/${TAG}_cuda python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-roofline > gpurun_out/${TAG}_ncu4.log 2>&1
for r in gemm256 gemm512 cuda; do
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_$r.ncu-rep.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$r.ncu-rep
done
tail -3 gpurun_out/${TAG}_tests.log
