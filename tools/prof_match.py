"""Profiling driver: a few 16k x 16k tensor-core matcher calls on device-resident data (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, MATCH_TENSOR, MATCH_TENSOR_FP16, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ALGO = MATCH_TENSOR if (len(sys.argv) > 3 and sys.argv[3] == "bf16x3") else MATCH_TENSOR_FP16
A, B, _ = synth.make_descriptors(n, n, 64, seed=2)
ctx = Context(0)
ctx.set_profiling(True)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
ts = []
for _ in range(reps):
    m = ctx.match_two_image(dA, dB, 0.3, algo=ALGO)
    ts.append(ctx.kernel_ms(0))
st = ctx.match_stats()
print("matches", len(m), "kernel ms", ts, "fallback rows", st.n_fallback_rows, "max_rel_err", st.max_rel_err)
