"""bench.py's host-side helpers (no GPU): the roofline traffic figure is read from the committed ncu summary, the
algorithmic work count follows SURVEY §8(d), and both arms describe the workload with the same config dict."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench_mod"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_traffic_comes_from_the_committed_profile():
    b = _bench()
    traffic, src = b.traffic_from_profile()
    assert isinstance(traffic, int) and traffic > 0
    assert src.startswith("profiles/") and "_ncu_mma_planes_full.txt" in src and "MB read" in src
    # the figure is the sum of the two dram counters of the FIRST launch block of that file
    path = os.path.join(ROOT, src.split(" ")[0])
    rd = wr = None
    for line in open(path):
        if "dram__bytes_read.sum" in line and rd is None:
            rd = float(line.split()[-2])
        if "dram__bytes_write.sum" in line and wr is None:
            wr = float(line.split()[-2])
    assert abs(traffic - (rd + wr) * 1e6) <= 1.0


def test_algorithmic_work_and_config_dicts():
    b = _bench()
    from clrsdp import solver
    bi = solver.BlockInfo(J=2, n_y=8, m=[1, 1], L=[1, 1], n_samples=[16, 16], Y_blocksizes=[[4], [4]], dim_S=[16, 16],
                          ranks=[[[1] * 16], [[1] * 16]])
    macs, pmac = b.algorithmic_int8_macs(bi, 256)
    assert pmac > 0 and macs == pmac * (32 * 33 // 2)          # s(s+1)/2 digit pairs with s = p/8
    assert b.config_dict("cfg3", 1) == b.config_dict("cfg3", 1)
    assert b.config_dict("cfg3", 8)["clusters_total"] == 8 * b.WORKLOADS["cfg3"]["J_per_gpu"]
    assert set(b.WORKLOADS) == {"cfg3", "cfg5shard"} and b.WORKLOADS["cfg5shard"]["prec"] == 512
