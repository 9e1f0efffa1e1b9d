"""One-off check at BASELINE sweep sizes: the tensor-core path against the exact SIMT kernel, full kNN tables bit for bit
(too slow for the test suite: the SIMT kernel needs ~40 ms at 64k x 64k)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, synth, MATCH_TENSOR, MATCH_SIMT_EXACT
ctx = Context(0)
for (nq, nt, seed) in ((65536, 65536, 1), (50001, 63999, 2)):
    A, B, _ = synth.make_descriptors(nq, nt, 64, seed=seed)
    B[100:140] = B[99]            # forty duplicates of one train row: more near-ties than a candidate list holds
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    t = ctx.match_two_image(a, b, 0.3, algo=MATCH_TENSOR, want_knn=True)
    st = ctx.match_stats()
    s = ctx.match_two_image(a, b, 0.3, algo=MATCH_SIMT_EXACT, want_knn=True)
    torch.cuda.synchronize()
    same_idx = torch.equal(t.knn_idx, s.knn_idx)
    same_d = torch.equal(t.knn_dist.view(torch.int32), s.knn_dist.view(torch.int32))
    print(nq, nt, "knn idx equal", same_idx, "knn dist bit-equal", same_d, "matches", len(t), len(s), "fallback rows", st.n_fallback_rows, "max_rel_err", st.max_rel_err)
