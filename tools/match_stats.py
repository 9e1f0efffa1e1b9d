import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, MATCH_TENSOR, MATCH_TENSOR_FP16, synth
ctx = Context(0)
ALGO = MATCH_TENSOR if (len(sys.argv) > 1 and sys.argv[1] == "bf16x3") else MATCH_TENSOR_FP16
for n in (16384, 32768, 65536):
    A, B, truth = synth.make_descriptors(n, n, 64, seed=n)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    m = ctx.match_two_image(dA, dB, 0.3, algo=ALGO)
    st = ctx.match_stats()
    print(n, "fallback rows", st.n_fallback_rows, "max_rel_err", st.max_rel_err, "tiles", st.n_tiles)
