"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Each rank solves its shard of a clustered SDP through NCCL-coupled handles AND the full problem on its own GPU
with an uncoupled handle; the log rows (mu, alpha, objectives) and the rank's slice of x, y must agree to
2^-(p-16) relative, and rank 0 also runs the CPU oracle on the whole problem and holds the N-GPU result to it (the sums over clusters are grouped differently, so the results are not bit-identical
between 1 and N GPUs, but they are bit-identical across the ranks of one run)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))
import numpy as np
import torch
import torch.distributed as dist

from clrsdp import instances, solver
from clrsdp.wire import rel_err_bits


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prec, Jloc = 256, 3
    kw = dict(delta=8, K=12, n_y=7, prec=prec, seed=17)
    full_c, full_b, _ = instances.synthetic_clustered_sdp(J=Jloc * world, **kw)
    my_c, my_b, _ = instances.synthetic_clustered_sdp(J=Jloc, j_offset=rank * Jloc, j_total=Jloc * world, **kw)
    assert np.array_equal(my_b.limb, full_b.limb)
    hf = solver.product_handle(prec, local)                 # the whole problem on this GPU
    bif = solver.get_block_info(full_c)
    solver.load_problem(hf, full_c, full_b, bif)
    hs = solver.product_handle(prec, local)                 # this rank's shard, coupled through NCCL
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        f = hs.lib.clrsdp_comm_unique_id
        f.argtypes = [ctypes.POINTER(ctypes.c_uint8)]
        assert f(buf) == 0
        uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    hs.comm_init(world, rank, bytes(uid.cpu().tolist()))
    bis = solver.get_block_info(my_c)
    solver.load_problem(hs, my_c, my_b, bis)
    for h in (hf, hs):
        h.set_params(solver.real_params(h.nlimb))
        h.init_point()
        h.prepare()
    ok = True
    nsl = sum(bis.dim_S)
    for it in range(4):
        rf, rs = hf.iterate(), hs.iterate()
        assert rf.status == 0 and rs.status == 0
        for k in ("mu", "alpha_p", "alpha_d", "beta_c", "p_obj_new", "d_obj_new", "primal_err_new", "dual_err_new"):
            a, b = getattr(rs, k), getattr(rf, k)
            # the error norms sit at rounding level (~2^-(p-40)) once a full step was taken: noise, not signal
            noise = 2.0 ** -(prec - 64) if k.endswith("err_new") else 1e-300
            if not np.isclose(a, b, rtol=1e-13, atol=noise):
                ok = False
                print(f"[rank {rank}] iter {it+1} {k}: sharded {a!r} vs single {b!r}", flush=True)
        xf, xs = hf.fetch("x"), hs.fetch("x")
        bits_x = rel_err_bits(xs, xf.take(range(rank * nsl, (rank + 1) * nsl)))
        bits_y = rel_err_bits(hs.fetch("y"), hf.fetch("y"))
        bits_Q = rel_err_bits(hs.fetch("Q"), hf.fetch("Q"))
        print(f"[rank {rank}] iter {it+1}: x {bits_x:.1f} y {bits_y:.1f} Q {bits_Q:.1f} bits; alpha_d {rs.alpha_d:.15e}", flush=True)
        ok = ok and min(bits_x, bits_y, bits_Q) >= prec - 16
    # the N-GPU result against the ORACLE run on the whole problem (rank 0 runs it; same 4 iterations)
    if rank == 0:
        from oracle.ref import oracle_handle
        ho = oracle_handle(prec, os.cpu_count() or 1)
        solver.load_problem(ho, full_c, full_b, bif)
        ho.set_params(solver.real_params(ho.nlimb))
        ho.init_point()
        ho.prepare()
        for it in range(4):
            ro = ho.iterate()
            assert ro.status == 0
        for k in ("mu", "alpha_p", "alpha_d", "beta_c", "p_obj_new", "d_obj_new"):
            a, b = getattr(rs, k), getattr(ro, k)
            if not np.isclose(a, b, rtol=1e-13, atol=1e-300):
                ok = False
                print(f"[rank 0] {k}: {world}-GPU {a!r} vs oracle {b!r}", flush=True)
        bo_x = rel_err_bits(hs.fetch("x"), ho.fetch("x").take(range(0, nsl)))
        bo_y = rel_err_bits(hs.fetch("y"), ho.fetch("y"))
        bo_Q = rel_err_bits(hs.fetch("Q"), ho.fetch("Q"))
        print(f"[rank 0] after 4 iterations vs the oracle: x {bo_x:.1f} y {bo_y:.1f} Q {bo_Q:.1f} bits", flush=True)
        ok = ok and min(bo_x, bo_y, bo_Q) >= prec - 16
    # replicated quantities are bit-identical across ranks
    y = hs.fetch("y")
    t = torch.tensor(y.limb.astype(np.int64).sum(axis=0) + y.exp, device="cuda")
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    ok = ok and all(torch.equal(lst[0], v) for v in lst)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(flag) else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
