"""Host-time vs device-time of sba_pair_rotation (device-resident), plus per-API cost estimate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, synth
import bench
ctx = Context(0)
pool = bench.make_pool(3, 1000)
dev = torch.device("cuda", 0)
res = []
for p in pool:
    res.append({k: torch.from_numpy(np.ascontiguousarray(p[s])).to(dev) for k, s in [("im1", "im1"), ("im2", "im2"), ("desc1", "desc1"), ("desc2", "desc2"), ("key1", "key1_xy"), ("key2", "key2_xy")]})
run = bench.PairRunner(ctx)
for k in range(5): run.run(res[k % 3])
torch.cuda.synchronize()
for want in (True, False):
    t0 = time.perf_counter()
    for k in range(50):
        d = res[k % 3]
        ctx.pair_rotation(d["im1"], d["im2"], d["desc1"], d["desc2"], d["key1"], d["key2"], bench.CS, want_matches=want)
    torch.cuda.synchronize()
    print("want_matches", want, "wall per pair (us):", (time.perf_counter() - t0) / 50 * 1e6)
# no images (matcher + BA only)
t0 = time.perf_counter()
for k in range(50):
    d = res[k % 3]
    ctx.pair_rotation(None, None, d["desc1"], d["desc2"], d["key1"], d["key2"], bench.CS, w=bench.W, h=bench.H, want_matches=False)
torch.cuda.synchronize()
print("no remap: wall per pair (us):", (time.perf_counter() - t0) / 50 * 1e6)
# matcher only
t0 = time.perf_counter()
for k in range(50):
    d = res[k % 3]
    ctx.match_two_image(d["desc1"], d["desc2"], 0.3)
torch.cuda.synchronize()
print("match_two_image only: wall (us):", (time.perf_counter() - t0) / 50 * 1e6)
import ctypes as C
b1 = torch.randn(8192, 4, device=dev); b2 = torch.randn(8192, 4, device=dev)
t0 = time.perf_counter()
for k in range(50):
    pr = ctx.ba_problem(b1, b2); r, s = pr.solve(np.zeros((1, 3))); pr.close()
torch.cuda.synchronize()
print("ba create+solve+destroy (8192 obs): wall (us):", (time.perf_counter() - t0) / 50 * 1e6, "iters", s.iterations, "evals", s.evaluations)
