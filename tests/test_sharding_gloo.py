"""N > 1 host logic on the CPU (world_size 2, gloo): the cluster sharding that bench.py / multi_gpu_check.py use.

Each rank generates only ITS clusters of the global instance (instances.synthetic_clustered_sdp with j_offset /
j_total), runs the first iteration of the CPU oracle on that shard and contributes its partial Q = sum_j W_j^T W_j.
The ranks exchange the partials with an all-gather and combine them in rank order — the same exchange pattern as the
NCCL path of the library (csrc/comm.cuh: all-gather + rank-ordered combine) — and rank 0 checks
  * the shards tile the global instance exactly (same b on every rank, B_j / c_j / V of cluster j identical to the
    single-process generation),
  * the combined Q equals the Q of the oracle run on the whole problem to 2^-(p-16) (sum(Q), MPMP.jl:1494),
  * the combined value is identical on both ranks (order-fixed combine => replicated state stays bit-identical).
The oracle is the checker here, as everywhere in tests/."""
import os
import socket
import sys
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

PREC, JLOC, KW = 128, 2, dict(delta=3, K=4, n_y=3, prec=128, seed=5)


def _first_iteration_Q(cons, b):
    from clrsdp import solver
    from oracle.ref import oracle_handle
    h = oracle_handle(PREC, 1)
    solver.load_problem(h, cons, b, solver.get_block_info(cons))
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    r = h.iterate()
    assert r.status == 0
    return h.fetch("Q")


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from clrsdp import instances
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        my_c, my_b, _ = instances.synthetic_clustered_sdp(J=JLOC, j_offset=rank * JLOC, j_total=JLOC * world, **KW)
        q = _first_iteration_Q(my_c, my_b)
        mine = dict(rank=rank, q=[q.get_int(i) for i in range(q.n)], b=my_b.limb.tobytes(),
                    B=[c.B.limb.tobytes() for c in my_c], c=[c.c.limb.tobytes() for c in my_c])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # rank-ordered combine of the partial sums (exact rationals stand in for the fixed-point lanes)
        comb = [sum((Fraction(g["q"][i][0]) * Fraction(2) ** g["q"][i][1] for g in sorted(gathered, key=lambda g: g["rank"])),
                    Fraction(0)) for i in range(q.n)]
        digest = [str(v) for v in comb]
        all_digests = [None] * world
        dist.all_gather_object(all_digests, digest)
        assert all(d == all_digests[0] for d in all_digests)
        if rank == 0:
            full_c, full_b, _ = instances.synthetic_clustered_sdp(J=JLOC * world, **KW)
            for g in gathered:
                assert g["b"] == full_b.limb.tobytes()
                for jl in range(JLOC):
                    assert g["B"][jl] == full_c[g["rank"] * JLOC + jl].B.limb.tobytes()
                    assert g["c"][jl] == full_c[g["rank"] * JLOC + jl].c.limb.tobytes()
            qf = _first_iteration_Q(full_c, full_b)
            scale = max(abs(qf.to_fraction(i)) for i in range(qf.n))
            worst = max(abs(comb[i] - qf.to_fraction(i)) for i in range(qf.n))
            assert worst <= scale * Fraction(1, 2 ** (PREC - 16)), float(worst / scale)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_cluster_sharding_combines_to_the_single_process_Q(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(tmp_path, "ok"))
