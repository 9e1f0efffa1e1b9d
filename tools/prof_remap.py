"""Remap throughput: batched equi2cube on device-resident frames (HBM roofline) -- also the ncu driver."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context

w, h, cs = 3840, 1920, 960
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = Context(0)
ctx.set_profiling(True)
g = torch.Generator(device="cuda"); g.manual_seed(1)
ims = torch.randint(0, 256, (nb, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
out = torch.empty((nb, cs, 6 * cs, 3), dtype=torch.uint8, device="cuda")
if len(sys.argv) > 4:
    ctx.set_remap_kernel(int(sys.argv[4]))      # 0 = plan's own choice, 1 = direct, 2 = tiled
mode = sys.argv[3] if len(sys.argv) > 3 else "cube"     # "cube": equi2cube strips; "bands": the 4 spherical_surf bands
ts = []
for _ in range(reps):
    if mode == "bands":
        ctx.spherical_crops(ims)
    else:
        ctx.equi2cube(ims, cs, out=out)
    ts.append(ctx.kernel_ms(1))
ms = float(np.median(ts[2:]))
px = nb * (4 * (h // 4) * w if mode == "bands" else cs * 6 * cs)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
alg = px * 6 + (nb * 0)  # 3 B gathered + 3 B written per output pixel (the 4 B/px index table is shared by all frames and L2-resident)
print(json.dumps({"plan": ctx.remap_plan_info(w, h, cs) if mode == "cube" else None, "mode": mode, "frames": nb, "kernel_ms": ms, "out_pixels": px, "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "peak_gbs": peak,
                  "frac": alg / ms / 1e6 / peak, "with_lut_bytes_gbs": (alg + px * 4) / ms / 1e6, "mpix_per_s": px / ms / 1e3}))
