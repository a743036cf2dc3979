"""Multi-GPU parity inside `pytest -m gpu`: launches tests/multi_gpu_check.py under torchrun on 2, 4 and 8 GPUs of this
box (skipped for the counts the box does not have). The check compares the N-GPU sharded solve with the single-GPU
solve AND with the CPU oracle on the whole problem, and the replicated state across ranks bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return sum(1 for ln in out.stdout.splitlines() if ln.startswith("GPU ")) if out.returncode == 0 else 0
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_solve_matches_single_gpu_and_oracle(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0 and "MULTI_GPU_PARITY OK" in out.stdout, tail


def test_single_process_multi_gpu_handle_matches_single_gpu_and_oracle():
    """clrsdp_create_multi: ONE handle, two GPUs, one process (what a Julia `ccall` front end uses). The whole problem
    goes in through the ordinary calls; the library partitions the clusters by weight (F16), routes the data and runs
    the iteration on both devices. Three iterations against the oracle on the general structure (field by field,
    global indices), bit-identical log rows against nothing less than the single-GPU handle's tolerance."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from clrsdp import instances, solver
    from oracle.ref import oracle_handle
    from test_gpu_solver import GENERAL_SPEC, compare_iteration
    prec = 256
    cons, b = instances.random_structured_sdp(GENERAL_SPEC, n_y=4, prec=prec)
    bi = solver.get_block_info(cons)
    hm, ho = solver.product_handle(prec, [0, 1]), oracle_handle(prec, 8)
    for h in (hm, ho):
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb))
        h.init_point()
        h.prepare()
    owner = hm.cluster_owner(bi.J)
    assert set(owner) == {0, 1}
    for it in range(3):
        rm, ro = hm.iterate(), ho.iterate()
        assert rm.status == 0 and ro.status == 0
        compare_iteration(hm, ho, bi, prec, prec - 16 - 8)     # (no p + 64 arbiter here: 8 bits of conditioning slack)
    # warm start through the multi handle: download (global order) -> upload -> same next iteration
    n_x, n_X = sum(bi.dim_S), sum(s * s for row in bi.Y_blocksizes for s in row)
    pt = hm.download_point(n_x, n_X, bi.n_y)
    r1 = hm.iterate()
    h2 = solver.product_handle(prec, [0, 1])
    solver.load_problem(h2, cons, b, bi)
    h2.set_params(solver.real_params(h2.nlimb))
    h2.upload_point(*pt)
    h2.prepare()
    r2 = h2.iterate()
    assert (r1.mu, r1.alpha_p, r1.alpha_d) == (r2.mu, r2.alpha_p, r2.alpha_d)


def test_sphere_packing_d12_sharded_over_two_gpus_matches_one_gpu():
    """Sphere packing d = 12 (clusters of very different weight: dim_S 2 x 169, 3 x 25-ish, 2 x 1) at 512 bits through a
    two-GPU handle: same iteration count as the single-GPU solve and the oracle's golden run (91), same objective."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import mpmath
    from clrsdp import instances, solver
    prec = 512
    solver.set_precision(prec)
    try:
        cons, b, _ = instances.sphere_packing_2point(n=3, d=12, prec=prec)
        bi = solver.get_block_info(cons)
        kw = dict(omega_p=100, omega_d=100, verbose=False, return_info=True)
        o1, r1 = solver.solverank1sdp(cons, b, bi, **kw)
        o2, r2 = solver.solverank1sdp(cons, b, bi, handle=solver.product_handle(prec, [0, 1]), **kw)
        with open(os.path.join(ROOT, "tests", "golden", "sphere_packing_512.json")) as f:
            g = [c for c in json.load(f)["cases"] if c["d"] == 12][0]
        assert len(r1) == len(r2) == g["iterations"] and r2[-1].terminate == 3
        with mpmath.workprec(prec):
            tol = mpmath.mpf(2) ** -128
            assert abs(o2[8] - o1[8]) <= tol and abs(o2[9] - o1[9]) <= tol
            assert abs(o2[8] - mpmath.mpf(g["primal_obj"])) <= tol
        from clrsdp.wire import rel_err_bits
        # x itself: the two runs differ in the grouping of the sum over clusters in Q (rounding level), amplified by
        # cond(S') - 2^240 and growing over the last iterations - on directions the objective does not see. Measured
        # agreement of x: 108 bits (objective: 128+ bits, identical iteration counts); the bar is twice the duality-gap
        # threshold (1e-15 ~ 50 bits).
        assert rel_err_bits(o2[0], o1[0]) > 100
    finally:
        solver.set_precision(256)
