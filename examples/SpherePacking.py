"""The reference's example (examples/SpherePacking.jl) on the B200 path: the Cohn-Elkies / de Laat-Oliveira-Vallentin
bound on the density of packings of spheres of N = 2 sizes in R^n, through the host mirror of the reference's
front end (`clrsdp.instances.sphere_packing_2point` restates ex:28-110; `clrsdp.solver.solverank1sdp` is the drop-in for
MPMP.jl:595-1025 and runs on the GPU — there is no CPU fallback).

    python examples/SpherePacking.py            # n = 3, d = 8 like ex:122: bound 0.8150097064...
    python examples/SpherePacking.py --d 16     # 0.8135955...
    python examples/SpherePacking.py --file-path /tmp/sdp --write-only   # the example's write_files / write_only (ex:95-98, :107)

Like the reference example (ex:29-31, 117-119) this runs at 512 bits: the Schur complements of this programme have
condition numbers of 2^126 (d = 8) to 2^185 (d = 16) at the first iteration already. The known values the
example quotes (ex:125-126): the bound is at least 0.793 and about 0.813 for high degree.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))
import mpmath  # noqa: E402

from clrsdp import instances, solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3, help="dimension (ex:122)")
    ap.add_argument("--d", type=int, default=8, help="polynomial degree parameter (ex:122)")
    ap.add_argument("--prec", type=int, default=512, help="working precision in bits (the example forces 512)")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--file-path", default="", help="write the constraints as an SDPB input directory first (ex:95-98)")
    ap.add_argument("--write-only", action="store_true", help="do not solve (the example's write_only, ex:107-115)")
    args = ap.parse_args()
    solver.set_precision(args.prec)                                   # setprecision(BigFloat, 512), ex:29-31
    cons, b, info = instances.sphere_packing_2point(n=args.n, d=args.d, prec=args.prec)
    blockinfo = solver.get_block_info(cons)                           # MPMP.jl:516-560
    if args.file_path:                                                # write_files(file_path, constraints, blockinfo, b), ex:97
        from clrsdp import sdpb_io
        sdpb_io.write_sdpb(args.file_path, cons, blockinfo, b)
        print("SDPB input directory written to", args.file_path)
    if args.write_only:
        return True
    out = solver.solverank1sdp(cons, b, blockinfo, omega_p=info["omega"], omega_d=info["omega"],
                               verbose=not args.quiet)                # ex:107-113
    x, X, y, Y, P, p, d, dual_gap, primal_obj, dual_obj, time_total = out
    with mpmath.workprec(args.prec):
        print("density bound (-primal objective):", mpmath.nstr(-primal_obj, 30))   # slots 9/10, not the time
        print("duality gap:", mpmath.nstr(dual_gap, 5), " time:", f"{time_total:.2f} s")


if __name__ == "__main__":
    main()
