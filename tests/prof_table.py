"""print the per-kernel CUDA-event table written by bench.py --profile-out"""
import json
import sys
d = json.load(open(sys.argv[1]))
tot = 0
for k, v in sorted(d["per_kernel"].items(), key=lambda kv: -kv[1]["ms"])[: int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f"{k:44s} {v['ms']/d['steps']:9.3f} ms/step {v['launches']//d['steps']:4d} launches/step")
print("sum of kernels per step:", sum(v["ms"] for v in d["per_kernel"].values()) / d["steps"], "ms; device seconds/step:", d["dev_seconds"] / d["steps"])
if "phase_ms_per_step" in d:
    print("phases (ms/step; decomposition = schur + chol_S + LinvB + Q + chol_Q, predictor/corrector = Z + rhs_x + system + dX + dY):")
    for k, v in d["phase_ms_per_step"].items():
        print(f"  {k:20s} {v:8.3f}")
