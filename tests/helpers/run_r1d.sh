mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -k "sphere or config4" 2>&1 | tail -60) > gpurun_out/r1d_newtests.log 2>&1
echo done
