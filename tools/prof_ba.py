"""Profiling driver: fused BA evaluation on a device-generated problem (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context
from tools.bench_ba import bearings_on_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
n_cam = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mat = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
ctx = Context(0)
b1, b2, cam, r_true = bearings_on_device(n, n_cam, dev, 5)
prob = ctx.ba_problem(b1, b2, cam if n_cam > 1 else None, n_cam)
ms = prob.eval_timed(r_true + 0.02, materialise=bool(mat), iters=5)
bpo = 80 if mat else 32
print(f"n={n} n_cam={n_cam} mat={mat}: {ms:.4f} ms  {n * bpo / ms / 1e6:.0f} GB/s")
