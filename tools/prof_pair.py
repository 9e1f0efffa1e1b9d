"""Profiling driver: a few C2-sized pairs through sba_pair_rotation on device-resident inputs (for ncu -k regex:...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, synth
W, H, CS, N = 3840, 1920, 960, 16384
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = Context(0)
pair = synth.make_pair(N, N, cs=CS, seed=5)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(im1=synth.make_erp_image(W, H, 1), im2=synth.make_erp_image(W, H, 2), d1=pair["desc1"],
     d2=pair["desc2"], k1=pair["key1_xy"], k2=pair["key2_xy"]).items()}
ctx.set_profiling(True)
for _ in range(reps):
    res, m, _ = ctx.pair_rotation(d["im1"], d["im2"], d["d1"], d["d2"], d["k1"], d["k2"], CS)
    print("matches", res.n_matches, "iterations", res.lm_iterations, "kernel ms: match %.4f remap %.4f ba %.4f" % (ctx.kernel_ms(0), ctx.kernel_ms(1), ctx.kernel_ms(2)))
