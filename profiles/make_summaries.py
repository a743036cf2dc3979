"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the text summaries committed here.
    python profiles/make_summaries.py r1i
reads gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum --csv), gpurun_out/<tag>_mma.ncu-rep and
gpurun_out/<tag>_others.ncu-rep (ncu --set full) and writes profiles/<tag>_*.txt / .csv."""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list ----
lines = [l for l in open(os.path.join(G, f"{tag}_launches.csv")) if not l.startswith("==")]
agg, tot, n = collections.OrderedDict(), 0.0, 0
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    ms = v / 1e6 if r["Metric Unit"] in ("ns", "nsecond") else (v / 1e3 if r["Metric Unit"] in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    tot += ms
    n += 1
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c {n}  (CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline [--no-second-roofline])",
       f"# {tag}: {n} consecutive launches = about two IPM iterations of BASELINE config 3, launched directly (the default bench replays",
       "# the same kernels from a CUDA graph; two streams are serialised by ncu).",
       "# Times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event shares (roofline.kernel_ms_per_step), not absolutes.",
       f"# total {tot:.3f} ms over {n} launches", f"{'kernel':<34}{'launches':>9}{'ms':>10}{'share':>8}"]
for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"{k:<34}{c:>9}{ms:>10.3f}{100 * ms / tot:>7.1f}%")
open(os.path.join(P, f"{tag}_ncu_launches_cfg3_summary.txt"), "w").write("\n".join(out) + "\n")
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{tag}_ncu_launches_cfg3.csv"))

# ---- full captures ----
WANT = ["Grid Size", "Block Size", "launch__shared_mem_per_block_dynamic", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def dump(rep, header, dst):
    # (round 2: the reports exceed what gpurun copies back, so the box exports `ncu -i X.ncu-rep --page raw --csv` to
    # X.ncu-rep.csv and deletes the report; either form is accepted here)
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        raw = open(rep + ".csv").read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = list(header) + [""]
    for r in data:
        lines.append("launch  " + r[idx["Kernel Name"]].split("(")[0][:90])
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                lines.append(f"  {w:<82}{r[idx[w]]:>16} {units[idx[w]]}")
    open(dst, "w").write("\n".join(lines) + "\n")


if os.path.exists(os.path.join(G, f"{tag}_gemm256.ncu-rep.csv")):   # round 2: single products, first launch = the dominant shape
    dump(os.path.join(G, f"{tag}_gemm256.ncu-rep"),
         ["# ncu --set full --clock-control none --cache-control none -k regex:mma_planes -c 20   (python tests/gpu_micro_gemm.py 256)",
          f"# {tag}: the sliced-GEMM tensor-core kernel on single products at 256 bit (T = 34 digit planes), 4 launches per shape, in this order:",
          "#   batch 64 x (64x64x64) [the dominant product of BASELINE config 3; fused carry], batch 128 x (32x32x32), 1 x (256x256x8192)",
          "#   [Q-shaped: K-split, unfused + carry launch], batch 64 x (256x128x128), batch 64 x (128x128x64).",
          "# --cache-control none: the caches are NOT flushed between the replay passes (the operands of a product come straight from the",
          "# slicing kernel through L2 in the solver as well).",
          f"# full report: gpurun_out/{tag}_gemm256.ncu-rep (scratch, not committed)"],
         os.path.join(P, f"{tag}_ncu_mma_planes_full.txt"))
if os.path.exists(os.path.join(G, f"{tag}_gemm512.ncu-rep.csv")):
    dump(os.path.join(G, f"{tag}_gemm512.ncu-rep"),
         ["# ncu --set full --clock-control none --cache-control none -k regex:mma_planes -c 16   (python tests/gpu_micro_gemm.py 512 cfg5)",
          f"# {tag}: the same kernel at the sizes of BASELINE config 5 (512 bit, T = 66 digit planes), 4 launches per shape, in this order:",
          "#   batch 64 x (128x128x128), batch 64 x (256x256x128), batch 16 x (1024x256x256), 1 x (1024x1024x16384) [Q: symmetric, K-split].",
          f"# full report: gpurun_out/{tag}_gemm512.ncu-rep (scratch, not committed)"],
         os.path.join(P, f"{tag}_ncu_mma_planes_cfg5_full.txt"))
if os.path.exists(os.path.join(G, f"{tag}_cuda.ncu-rep.csv")):
    dump(os.path.join(G, f"{tag}_cuda.ncu-rep"),
         ['# ncu --set full --clock-control none -k regex:"slice_rows|panel_factor|gemv|small_gemm|lambda_min" -s 300 -c 24   (CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-roofline)',
          f"# {tag}: the CUDA-core kernels around the sliced GEMM (BASELINE config 3): occupancy, DRAM traffic, issue-stall shares.",
          "# (cold caches between replay passes and streams serialised: DRAM figures are upper bounds)",
          f"# full report: gpurun_out/{tag}_cuda.ncu-rep (scratch, not committed)"],
         os.path.join(P, f"{tag}_ncu_cuda_core_kernels_full.txt"))
if os.path.exists(os.path.join(G, f"{tag}_mma.ncu-rep")):
    dump(os.path.join(G, f"{tag}_mma.ncu-rep"),
         ["# ncu --set full --clock-control none --import-source on -k regex:mma_planes -s 90 -c 6   (CLRSDP_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline)",
          f"# {tag}: six consecutive mma_planes_kernel launches of an IPM iteration of BASELINE config 3 (256 bit, T = 34 digit planes).",
          f"# full report: gpurun_out/{tag}_mma.ncu-rep (scratch, not committed)"],
         os.path.join(P, f"{tag}_ncu_mma_planes_full.txt"))
if os.path.exists(os.path.join(G, f"{tag}_others.ncu-rep")):
    dump(os.path.join(G, f"{tag}_others.ncu-rep"),
         ['# ncu --set full --clock-control none --import-source on -k regex:"carry_kernel|slice_rows|panel_factor" -s 200 -c 12   (same command)',
          f"# {tag}: the CUDA-core kernels around the sliced GEMM (BASELINE config 3): occupancy, DRAM traffic, issue-stall shares.",
          f"# full report: gpurun_out/{tag}_others.ncu-rep (scratch, not committed)"],
         os.path.join(P, f"{tag}_ncu_carry_slice_panel_full.txt"))
