#!/usr/bin/env python
"""bench.py — seconds per interior-point iteration of the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is ONE IPM iteration (one pass of MPMP.jl:754-953) on the synthetic clustered low-rank SDP of
BASELINE config 3 (64 clusters per GPU, rank-1 constraints, block 64, 128 samples, n_y = 256, 256-bit;
weak scaling: N GPUs hold 64*N clusters — at N = 8 this is the north star's 512-cluster instance).
`value` is the device time per iteration with the state resident in HBM; `e2e` is the same iteration
through the C ABI with the iterate (x, X, y, Y) living in HOST buffers: upload, iterate, download.
`--impl reference` times the CPU restatement of the reference (oracle/, MPFR) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE config 3 per GPU (the configuration the metric is quoted on; at N = 8 the north star's 512 clusters)
    "cfg3": dict(J_per_gpu=64, delta=64, K=128, n_y=256, prec=256, seed=20261018),
    # one GPU's share of BASELINE config 5 on 8 GPUs (512 clusters / 8, block 128, 512 bit): not the headline, run with
    # `--workload cfg5shard --no-cpu-baseline` to see the kernels at the sizes they were designed for
    "cfg5shard": dict(J_per_gpu=64, delta=128, K=256, n_y=1024, prec=512, seed=20261019),
}
WORKLOAD = dict(WORKLOADS["cfg3"])
METRIC = "sec/IPM iteration at 256-bit (Schur build+Cholesky), 1/2/4/8 B200 vs CPU"
UNIT = "s/iteration"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_tflops=d.get("bf16_tflops", 1590.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md): NVML when the binding is importable (a
    sample every ~5 ms), else nvidia-smi (one sample per ~60 ms process start)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index   # rows: (sm_mhz, sm_max_mhz, set of reasons)
        self.how = "nvidia-smi"

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        hd = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(hd, nv.NVML_CLOCK_SM))
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        self.how = "nvml"
        while not self.stop_flag:
            sm = float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(hd))
            self.rows.append((sm, mx, {n for b, n in bits if r & b}))
            time.sleep(0.005)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            self.how = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [v.strip() for v in out.split(",")]
                if len(c) >= 6:
                    self.rows.append((float(c[0]), float(c[1]), {self.NAMES[i] for i in range(4) if c[2 + i].lower().startswith("active")}))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted(set().union(*[r[2] for r in self.rows])) if self.rows else []
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows), how=self.how)


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant sliced-GEMM launch (batch-64 64x64x64 at 256 bit), read
    from the committed ncu summary (newest profiles/r*_ncu_mma_planes_full.txt) - never a literal in this file."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_mma_planes_full.txt")), key=os.path.getmtime)
    for path in reversed(files):
        rd = wr = None
        grid = None
        with open(path) as f:
            for line in f:
                m = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", line)
                if m and rd is None:
                    rd = float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
                m = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", line)
                if m and wr is None:
                    wr = float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
                m = re.search(r"Grid Size\s+\((\d+)", line)
                if m and grid is None:
                    grid = int(m.group(1))
                if rd is not None and wr is not None:
                    break
        if rd is not None and wr is not None:
            return int(rd + wr), f"{os.path.relpath(path, ROOT)} (first launch: grid {grid}, {rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB written)"
    return None, "no ncu summary under profiles/"


def tensor_pipe_from_profile(pattern):
    """sm__pipe_tensor_cycles_active (% of peak, per launch) of the tensor kernel from the newest committed ncu summary that
    matches `pattern` under profiles/ - the hardware's own view beside `roofline.frac`, which since the carry propagation
    was fused into the kernel counts the epilogue's CUDA-core work (carry, normalise, round, store) in its denominator."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
    if not files:
        return None
    vals, times = [], []
    with open(files[-1]) as f:
        for line in f:
            m = re.search(r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active\s+([0-9.]+)", line)
            if m:
                vals.append(round(float(m.group(1)), 1))
            m = re.search(r"gpu__time_duration\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                times.append(round(float(m.group(1)) * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3,
                                                        "nsecond": 1e-3}.get(m.group(2), 1.0), 1))
    # the captures hold 4 launches per product shape: one figure per shape
    per_shape = [vals[i] for i in range(0, len(vals), 4)]
    return dict(source=os.path.relpath(files[-1], ROOT), pct_per_product_shape=per_shape,
                us_per_launch=[times[i] for i in range(0, len(times), 4)])


def mma_roofline_of(name, steps, warmup, int8_peak_tops, device=0):
    """The sliced-GEMM roofline entry on another workload in the same run (one GPU): `steps` profiled iterations."""
    from clrsdp import solver
    saved = dict(WORKLOAD)
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[name])
    try:
        prec = WORKLOAD["prec"]
        cons, b, bi = build_problem(WORKLOAD["J_per_gpu"], 0, WORKLOAD["J_per_gpu"], prec)
        h = solver.product_handle(prec, device)
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb))
        h.init_point()
        h.prepare()
        dev = []
        for i in range(warmup + steps):
            r = h.iterate()
            if i >= warmup:
                dev.append(r.seconds)
        h.init_point()
        h.prepare()
        h.profile_reset(True)
        prof_s = 0.0
        for _ in range(steps):
            prof_s += h.iterate().seconds
        prof = h.profile_dump()
        h.profile_reset(False)
        mma = dict(ms=0.0, launches=0, work=0.0)
        for k, v in prof.items():
            if k.startswith("mma_planes"):
                for f in mma:
                    mma[f] += v[f]
        tops = (2.0 * mma["work"] / (mma["ms"] * 1e-3) / 1e12) if mma["ms"] > 0 else 0.0
        macs, pmac = algorithmic_int8_macs(bi, prec)
        return dict(workload=name, config=dict(WORKLOAD), bound="tensor", kernel="mma_planes_kernel", achieved=tops, peak=int8_peak_tops,
                    unit="TOP/s (int8)", frac=tops / int8_peak_tops if int8_peak_tops else None, launches=mma["launches"],
                    ms_per_step=float(np.mean(dev)) * 1e3, steps=steps, share_of_step=mma["ms"] * 1e-3 / prof_s if prof_s else None,
                    int8_mac_per_iter=macs, tensor_pipe_active_ncu=tensor_pipe_from_profile("r*_ncu_mma_planes_cfg5_full.txt"))
    finally:
        WORKLOAD.clear()
        WORKLOAD.update(saved)


def build_problem(J, j_offset, j_total, nl_prec):
    from clrsdp import instances, solver
    cons, b, info = instances.synthetic_clustered_sdp(J=J, delta=WORKLOAD["delta"], K=WORKLOAD["K"], n_y=WORKLOAD["n_y"],
                                                      prec=nl_prec, seed=WORKLOAD["seed"], j_offset=j_offset, j_total=j_total)
    return cons, b, solver.get_block_info(cons)


def algorithmic_int8_macs(bi, prec):
    """pMACs of the pure-GEMM phases actually executed per iteration x s(s+1)/2, s = p/8."""
    s = prec // 8
    per = s * (s + 1) // 2
    pmac = 0
    for j in range(bi.J):
        m, K, dimS = bi.m[j], bi.n_samples[j], bi.dim_S[j]
        for l in range(bi.L[j]):
            nb, dl = bi.Y_blocksizes[j][l], bi.delta[j][l]
            Nv = int(sum(bi.ranks[j][l]))
            pmac += 2 * (nb * nb * Nv + m * nb * Nv * Nv)          # pairings (X^-1 and Y)
            pmac += 2 * nb ** 3                                    # XY, dXdY
            pmac += nb ** 3                                        # X^-1 = V V^T
            pmac += 2 * (4 * nb ** 3)                              # Z and dY, two directions
            pmac += 3 * (m * (m + 1) // 2) * dl * dl * Nv          # V D V^T: residual + two directions
            pmac += 2 * nb * nb * Nv                               # Z V for trace_A, two directions
            pmac += 2 * (2 * nb ** 3)                              # L^-1 dM L^-T for X and Y
        pmac += dimS * dimS * bi.n_y                               # W = L^-1 B
        pmac += bi.n_y * (bi.n_y + 1) // 2 * dimS                  # Q = W^T W: symmetric, upper triangle computed (SYRK)
    return pmac * per, pmac


def config_dict(workload_name, world):
    """The `config` object of the JSON line: identical for the GPU arm and for `--impl reference`."""
    Jloc = WORKLOAD["J_per_gpu"]
    return dict(workload=("synthetic clustered low-rank SDP, BASELINE config 3 per GPU " if workload_name == "cfg3" else
                          "synthetic clustered low-rank SDP, one GPU's share (64 clusters) of BASELINE config 5 ") +
                         "(manufactured strictly feasible; iterations from omega*I)",
                clusters_total=world * Jloc, l2="working set > L2 (several hundred MB of arenas touched per "
                "iteration); no explicit flush", **WORKLOAD)


def _time_oracle(n_threads, J, iters, gemm_mode, discard=1):
    """Wall-clock seconds of `iters` FULL interior-point iterations of the oracle on J clusters (no sub-sampling, no
    extrapolation), after `discard` untimed ones: the first iteration starts from omega*I, whose zero off-diagonal
    entries make MPFR products cheaper than in any later iteration (the reference drops its first two, MPMP.jl:889)."""
    from clrsdp import solver
    from oracle.ref import oracle_handle
    cons, b, bi = build_problem(J, 0, J, WORKLOAD["prec"])
    if gemm_mode:
        os.environ["CLRSDP_REF_GEMM"] = gemm_mode      # read by the oracle when a handle is created
    else:
        os.environ.pop("CLRSDP_REF_GEMM", None)
    h = oracle_handle(WORKLOAD["prec"], n_threads)
    os.environ.pop("CLRSDP_REF_GEMM", None)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    per_iter = []
    for i in range(discard + iters):
        t0 = time.time()
        h.iterate()
        if i >= discard:
            per_iter.append(time.time() - t0)
    h.close()
    return per_iter


def cpu_baseline(n_threads, J=None, iters=1, both=False):
    """The oracle (CPU restatement of MPMP.jl, MPFR; the reference itself - Julia + Arb - cannot run in this image) timed
    on the host cores on the SAME instance as the GPU arm: all J clusters, whole iterations, nothing extrapolated.
    `value` is timed with the oracle's block fixed-point product (CLRSDP_REF_GEMM=fixed: exact integer dot products
    over mpn limbs, one rounding per entry - the way libarb's approx_mul works); the time with classical
    mpfr_fma triple loops (what the parity tests run) is reported beside it."""
    J = J or WORKLOAD["J_per_gpu"]
    t_fixed = _time_oracle(n_threads, J, iters, "fixed")
    out = dict(value=float(np.mean(t_fixed)), unit=UNIT, cores=n_threads, kind="port", clusters=J, iterations_timed=len(t_fixed),
               per_iteration_s=[round(t, 3) for t in t_fixed],
               sample=f"{len(t_fixed)} whole iteration(s) (after 1 untimed from omega*I) of the MPFR restatement (oracle/, block "
                      f"fixed-point GEMM over mpn limbs like libarb's approx_mul) on all {J} clusters of the workload, "
                      f"{n_threads} threads over the clusters / blocks / Q chunks like the reference's Threads.@threads; "
                      "the reference itself (Julia+Arb) cannot run in this image")
    if both:
        out["value_mpfr_fma_loops"] = float(np.mean(_time_oracle(n_threads, J, 1, None)))
    return out


def run_reference(args):
    """`--impl reference`: the reference's CPU path for the same metric and config. Every timed step is one WHOLE
    iteration on the GPU arm's instance (64 clusters per GPU x N GPUs) on all host cores; `steps` is the number of
    iterations actually timed (the run is capped at a few minutes whatever K is). Should the N-GPU instance not fit that
    cap on this host, fewer clusters are timed and `work_ratio` (GPU-arm work / timed work) says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    Jloc, world = WORKLOAD["J_per_gpu"], max(1, args.gpus)
    budget = 200.0
    t_begin = time.time()
    t64 = _time_oracle(n_threads, Jloc, 1, "fixed", discard=1)[0]      # also the warm-up (library load, page faults)
    J = Jloc * world
    if world > 1 and 2.5 * t64 * world > budget:                       # (1 discarded + >= 1 timed iteration must fit)
        J = Jloc * max(1, min(world, int(budget / (2.5 * t64))))
    if J == Jloc:
        reps = int(max(1, min(args.steps, (budget - (time.time() - t_begin)) / max(t64, 1e-3))))
        vals = [t64] + (_time_oracle(n_threads, J, reps - 1, "fixed") if reps > 1 else [])
    else:
        est = t64 * J / Jloc
        reps = int(max(1, min(args.steps, (budget - (time.time() - t_begin)) / est - 1)))
        vals = _time_oracle(n_threads, J, reps, "fixed")
    v = float(np.mean(vals))
    cb = dict(value=v, unit=UNIT, cores=n_threads, kind="port", clusters=J, iterations_timed=len(vals),
              per_iteration_s=[round(t, 3) for t in vals],
              sample=f"{len(vals)} whole iteration(s) of the MPFR restatement (oracle/, block fixed-point GEMM over mpn limbs like "
                     f"libarb's approx_mul) on {J} clusters, {n_threads} threads; nothing extrapolated")
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(vals), warmup=1,
                ms_per_step=v * 1e3, higher_is_better=False, scaling="weak", vs_baseline=None,
                dtype=f"u{WORKLOAD['prec']} fixed-limb float (int8 slices, int32 accumulate)", data="synthetic",
                impl="reference", steps_requested=args.steps, warmup_requested=args.warmup,
                config=config_dict(args.workload, world), work_ratio=(Jloc * world) / J,
                reference_dtype=f"f{WORKLOAD['prec']} (MPFR round-to-nearest; block fixed-point products)",
                cpu_baseline=cb, e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-fma", action="store_true", help="also time the oracle with classical mpfr_fma product loops")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here")
    ap.add_argument("--no-second-roofline", action="store_true",
                    help="skip the roofline entry of the cfg5shard workload (a few iterations at BASELINE config 5's sizes)")
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    if args.impl == "reference":
        return run_reference(args)

    from clrsdp import solver
    from clrsdp.wire import MpArray
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prec = WORKLOAD["prec"]
    Jloc = WORKLOAD["J_per_gpu"]
    cons, b, bi = build_problem(Jloc, rank * Jloc, world * Jloc, prec)
    h = solver.product_handle(prec, local_rank)
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            f = h.lib.clrsdp_comm_unique_id
            f.argtypes = [ctypes.POINTER(ctypes.c_uint8)]
            st = f(buf)
            if st != 0:
                raise RuntimeError("clrsdp_comm_unique_id failed")
            uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        h.comm_init(world, rank, bytes(uid.cpu().tolist()))
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing (per-kernel profiling OFF: it adds two event records per launch) ----
    # The solve converges in a few dozen iterations; every RESTART steps the iterate is put back to omega*I (the host
    # side of that, ~1 ms, is outside the device time that `value` reports and is subtracted from the launch count) so
    # that every timed step is a regular iteration, however large K is.
    RESTART = 20

    def restart():
        h.init_point()
        h.prepare()

    for _ in range(args.warmup):
        r = h.iterate()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_s = []
    launches = 0
    for i in range(args.steps):
        if i and i % RESTART == 0:
            restart()
        l0 = h.launch_count()
        r = h.iterate()            # blocks until the iteration's log row is back: device work is complete
        launches += h.launch_count() - l0
        dev_s.append(r.seconds)
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    dev_total = float(np.sum(dev_s))
    if dist is not None:
        import torch
        t = torch.tensor([dev_total, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total, wall = float(t[0]), float(t[1])
    sec_per_iter = dev_total / args.steps
    # ---- per-kernel pass: the same iterations again with CUDA events around every launch (roofline numbers) ----
    prof_s = 0.0
    prof_acc = {}
    phase_names = ["decomposition", "predictor", "corrector", "step_length", "factor_XY+X_inv", "R", "residuals", "schur",
                   "chol_S", "LinvB", "Q", "chol_Q", "Z", "rhs_x", "system", "dX", "dY"]
    phases = np.zeros(len(phase_names))
    restart()
    h.profile_reset(True)          # drop the restart's kernels from the table
    for i in range(args.steps):
        if i and i % RESTART == 0:
            saved = h.profile_dump()
            h.profile_reset(False)
            restart()
            h.profile_reset(True)
            for k, v in saved.items():
                keep = prof_acc.setdefault(k, dict(ms=0.0, launches=0, work=0.0))
                for f in keep:
                    keep[f] += v[f]
        rr = h.iterate()
        prof_s += rr.seconds
        phases += np.array(list(rr.timings)[:len(phase_names)])
    prof = h.profile_dump()
    for k, v in prof_acc.items():
        keep = prof.setdefault(k, dict(ms=0.0, launches=0, work=0.0))
        for f in keep:
            keep[f] += v[f]
    h.profile_reset(False)

    # ---- end to end: the iterate lives in host buffers between iterations ----
    n_x, n_X, n_y = int(sum(bi.dim_S)), int(sum(s * s for row in bi.Y_blocksizes for s in row)), bi.n_y
    state = h.download_point(n_x, n_X, n_y)
    h.pin(*state)   # the iterate's host buffers are page-locked once; every step DMAs from / into them
    bytes_per_num = 4 * h.nlimb + 8 + 1
    h2d = (n_x + n_y + 2 * n_X) * bytes_per_num
    d2h = h2d + ctypes.sizeof(type(r))
    state_out = h.download_point(n_x, n_X, n_y)
    h.pin(*state_out)              # (one-off allocation and page-locking of the result buffers: outside the timed region)
    e2e_parts = np.zeros(4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):    # every step: host iterate -> device, one iteration, new iterate -> host
        ta = time.perf_counter()
        h.upload_point(*state)
        tb = time.perf_counter()
        h.prepare()
        tc = time.perf_counter()
        r = h.iterate()
        td = time.perf_counter()
        state_out = h.download_point(n_x, n_X, n_y, out=state_out)
        te = time.perf_counter()
        e2e_parts += np.array([tb - ta, tc - tb, td - tc, te - td])
    barrier()
    e2e_wall = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0])
    sampler.join(timeout=2)

    if dist is not None:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return
    peaks = measured_peaks()
    macs, pmac = algorithmic_int8_macs(bi, prec)
    mma = dict(ms=0.0, launches=0, work=0.0)
    for k, v in prof.items():          # kernel names carry the GEMM shape: mma_planes_M.._N.._K.._b..
        if k.startswith("mma_planes"):
            for f in mma:
                mma[f] += v[f]
    # int8 dense tensor peak: not in MEASURED_PEAKS.json, so it is measured here, in this run, on this GPU
    # (clrsdp_measure_int8_peak: back-to-back tcgen05.mma kind::i8 128x256x32 on every SM, operands resident in smem)
    int8_peak_macs = h.measure_int8_peak()
    int8_peak_tops = 2.0 * int8_peak_macs / 1e12
    mma_tops = (2.0 * mma["work"] / (mma["ms"] * 1e-3) / 1e12) if mma["ms"] > 0 else 0.0
    agg = {}
    for k, v in prof.items():
        base = k.split("_M")[0] if k.startswith("mma_planes") else k
        a = agg.setdefault(base, dict(ms=0.0, launches=0))
        a["ms"] += v["ms"]
        a["launches"] += v["launches"]
    top = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]
    traffic, traffic_src = traffic_from_profile()
    roofline = dict(bound="tensor", kernel="mma_planes_kernel", achieved=mma_tops, peak=int8_peak_tops, unit="TOP/s (int8)",
                    frac=mma_tops / int8_peak_tops if int8_peak_tops else None,
                    traffic=traffic if args.workload == "cfg3" else None, traffic_source=traffic_src,
                    tensor_pipe_active_ncu=tensor_pipe_from_profile("r*_ncu_mma_planes_full.txt" if args.workload == "cfg3"
                                                                    else "r*_ncu_mma_planes_cfg5_full.txt"),
                    frac_note="algorithmic int8 MACs / event time of the whole tensor kernel / measured int8 peak; since round 2 "
                              "the kernel's time includes the fused carry / normalise / store epilogue (a separate launch in "
                              "round 1), so frac is not comparable with round 1's 0.13; tensor_pipe_active_ncu is the hardware counter",
                    peak_source="measured in this run (tcgen05.mma kind::i8 issue loop on all SMs); MEASURED_PEAKS.json has "
                                f"no int8 entry (its bf16 figure: {peaks['bf16_tflops']} TFLOP/s, {peaks['source']})",
                    achieved_note="ALGORITHMIC int8 ops (M*N*K*s(s+1)/2 MACs per product, s = p/8; no guard digits, no tile "
                                  "padding) of all sliced-GEMM launches of the step / their summed CUDA-event time",
                    launches=mma["launches"], ms_per_launch=mma["ms"] / max(1, mma["launches"]),
                    share_of_step=mma["ms"] * 1e-3 / prof_s if prof_s else None,
                    kernel_ms_per_step={k: round(v["ms"] / args.steps, 4) for k, v in top})
    line = dict(metric=METRIC, value=sec_per_iter, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec_per_iter * 1e3, higher_is_better=False, scaling="weak", vs_baseline=None,
                dtype=f"u{prec} fixed-limb float (int8 slices, int32 accumulate)", data="synthetic",
                config=config_dict(args.workload, world),
                wall_ms_per_step=wall / args.steps * 1e3,
                e2e=dict(value=e2e_wall / args.steps, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         parts_ms=dict(zip(["upload_point", "prepare", "iterate", "download_point"],
                                           [round(float(v) / args.steps * 1e3, 3) for v in e2e_parts])),
                         note="upload_point + prepare + iterate + download_point through the C ABI with host buffers"),
                gpu_launches=int(launches), clocks=sampler.summary(), roofline=roofline,
                algorithmic=dict(pmac_per_iter=pmac, int8_mac_per_iter=macs))
    if world == 1 and args.workload == "cfg3" and not args.no_second_roofline:
        # the sliced-GEMM kernel at the sizes it was designed for (one GPU's share of BASELINE config 5: block 128, K = 256,
        # n_y = 1024, 512 bit): a second roofline entry in the same driver-visible line
        try:
            del h
            line["roofline_cfg5shard"] = mma_roofline_of("cfg5shard", 4, 2, int8_peak_tops, local_rank)
        except Exception as e:
            line["roofline_cfg5shard"] = dict(error=str(e))
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1, both=args.cpu_fma)
        except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
            line["cpu_baseline"] = dict(error=str(e))
    if args.profile_out:
        with open(args.profile_out, "w") as f:
            json.dump(dict(per_kernel=prof, steps=args.steps, dev_seconds=prof_s,
                           phase_ms_per_step={k: float(v) / args.steps * 1e3 for k, v in zip(phase_names, phases)}), f, indent=1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
