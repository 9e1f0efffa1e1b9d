"""One pair at a time through sba_pair_rotation_begin/_end on device-resident inputs: where the wall time goes
(host time inside begin = enqueue of the 7 launches; time inside end = wait for the device + result read-back)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context
import bench
ctx = Context(0)
pool = bench.make_pool(3, 1000)
dev = torch.device("cuda", 0)
res = [{k: torch.from_numpy(np.ascontiguousarray(p[s])).to(dev) for k, s in [("im1", "im1"), ("im2", "im2"), ("desc1", "desc1"), ("desc2", "desc2"), ("key1", "key1_xy"), ("key2", "key2_xy")]} for p in pool]
run = bench.PairRunner(ctx)
for k in range(6): run.run(res[k % 3])
torch.cuda.synchronize()
n = 60
tb = te = 0.0
t_all0 = time.perf_counter()
for k in range(n):
    d = res[k % 3]
    t0 = time.perf_counter()
    call = run.begin(d)
    t1 = time.perf_counter()
    run.collect(call)
    t2 = time.perf_counter()
    tb += t1 - t0; te += t2 - t1
t_all = time.perf_counter() - t_all0
print(json.dumps({"pairs": n, "wall_us_per_pair": t_all / n * 1e6, "host_us_in_begin": tb / n * 1e6, "host_us_in_end": te / n * 1e6}))
