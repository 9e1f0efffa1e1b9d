"""Profiling driver: the initial-guess normal-matrix pass on a large match list (for ncu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spherical_bundle_adjuster_b200 import Context, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = Context(0)
b1, b2, r, t, _ = synth.make_two_view(n, seed=3)
rng = np.random.default_rng(0)
idx = np.stack([rng.permutation(n)[: n // 4] for _ in range(80)]).astype(np.int32)
for k in range(3):
    t0 = time.perf_counter()
    R, T, nc = ctx.initial_guess(b1, b2, idx)
    dt = time.perf_counter() - t0
print(f"n={n}: initial_guess {dt * 1e3:.2f} ms host wall (80 subsets of {n // 4}), R={R}, candidates={nc}")
