"""Per-kernel CUDA-event timings of the phase-level entry points (run on the GPU box):
    python tests/gpu_micro.py [prec]
Not a test: a measuring aid for the latency-bound kernels (panel factorisation, lambda_min, GEMM launches)."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
sys.path.insert(0, os.path.dirname(__file__))
import numpy as np  # noqa: E402

from clrsdp import solver  # noqa: E402
from clrsdp.wire import MpArray  # noqa: E402
from gpu_common import spd_batch  # noqa: E402


def table(h, title, reps):
    prof = h.profile_dump()
    print(f"== {title}")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:40s} {v['ms'] / reps:9.3f} ms/call {v['launches'] // reps:4d} launches  {v['ms'] / v['launches'] * 1e3:9.1f} us/launch")
    print(f"   total {sum(v['ms'] for v in prof.values()) / reps:.3f} ms")


def main():
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    only = sys.argv[2] if len(sys.argv) > 2 else "all"   # "chol": the factorisations only
    h = solver.product_handle(prec, 0)
    rng = random.Random(1)
    for batch, n in [(128, 64), (64, 128), (1, 256), (128, 32), (1, 32)]:
        A = spd_batch(rng, batch, n, h.nlimb)
        h.op_cholesky(batch, n, A)
        h.profile_reset(True)
        for _ in range(3):
            h.op_cholesky(batch, n, A)
        table(h, f"cholesky+inverse batch={batch} n={n}", 3)
        h.profile_reset(False)
    if only == "chol":
        return
    for batch, n in [(128, 64), (1, 64), (16, 128)]:
        mats = []
        nrng = np.random.default_rng(n)
        for _ in range(batch):
            G = nrng.uniform(-1, 1, size=(n, n))
            mats.append((G + G.T) / 2)
        A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
        h.op_lambda_min(batch, n, A)
        h.profile_reset(True)
        for _ in range(3):
            h.op_lambda_min(batch, n, A)
        table(h, f"lambda_min batch={batch} n={n}", 3)
        h.profile_reset(False)
    for batch, M, N, K in [(64, 64, 64, 64), (1, 32, 32, 32), (1, 256, 256, 8192), (64, 256, 128, 128)]:
        nrng = np.random.default_rng(1)
        A = MpArray.from_double(nrng.uniform(-1, 1, size=batch * M * K), h.nlimb)
        B = MpArray.from_double(nrng.uniform(-1, 1, size=batch * K * N), h.nlimb)
        h.op_gemm(batch, M, N, K, A, B)
        h.profile_reset(True)
        for _ in range(3):
            h.op_gemm(batch, M, N, K, A, B)
        table(h, f"gemm batch={batch} M={M} N={N} K={K}", 3)
        h.profile_reset(False)


if __name__ == "__main__":
    main()
