"""GPU parity for BASELINE config 3 (all-pairs matching of a frame sequence, pairs sharded across ranks): every pair of
every shard, matched with two library contexts in flight, equals the oracle bit for bit."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import Context, sharding, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prepared", [False, True])
def test_all_pairs_sharded_matches_oracle(prepared):
    import torch
    F, n = 5, 1100                                   # 10 pairs; 1100 x 1100 takes the tensor-core path
    rng = np.random.default_rng(3)
    base = synth.unit_rows(rng.standard_normal((n, 64)))
    frames = [synth.unit_rows(base[rng.permutation(n)] + 0.02 * rng.standard_normal((n, 64))).astype(np.float32) for _ in range(F)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
    if prepared:     # every frame handed over once (sba_descriptors_create), used by both contexts
        dev = [ctxs[k % 2].prepare_descriptors(f) for k, f in enumerate(frames)]
    else:
        dev = [torch.from_numpy(f).cuda() for f in frames]
    seen = []
    for rank in range(2):                            # the two shards of a world of 2, one after the other
        pairs = sharding.shard_pairs(F, rank, 2)
        calls = []
        for k, (i, j) in enumerate(pairs):
            if len(calls) == 2:                      # two in flight
                _check(*calls.pop(0), frames)
            calls.append((ctxs[k % 2].match_begin(dev[i], dev[j]), i, j))
        for c in calls:
            _check(*c, frames)
        seen += pairs
    assert seen == sharding.all_pairs(F)
    for c in ctxs:
        c.close()


def _check(call, i, j, frames):
    m = call.end()
    qi, ti, dd = oracle.match_two_image(frames[i], frames[j], 0.3)
    assert len(qi) > 100
    assert np.array_equal(m.query_idx.cpu().numpy(), qi) and np.array_equal(m.train_idx.cpu().numpy(), ti)
    assert np.array_equal(m.distance.cpu().numpy().view(np.uint32), dd.view(np.uint32))


@pytest.mark.parametrize("nq,nt,dim", [(1500, 1300, 64), (300, 200, 64), (700, 900, 128), (1, 2000, 64)])
def test_prepared_sets_give_identical_results(ctx, nq, nt, dim):
    """sba_knn2_ratio_prepared against sba_knn2_ratio and the oracle: tensor path, SIMT path, 128-d, ragged sizes."""
    A, B, _ = synth.make_descriptors(nq, nt, dim, seed=nq)
    pa, pb = ctx.prepare_descriptors(A), ctx.prepare_descriptors(B)
    assert len(pa) == nq and len(pb) == nt
    m = ctx.match_two_image(pa, pb, 0.3, want_knn=True)
    ref = ctx.match_two_image(A, B, 0.3, want_knn=True)
    qi, ti, dd = oracle.match_two_image(A, B, 0.3)
    assert np.array_equal(m.query_idx, qi) and np.array_equal(m.train_idx, ti) and np.array_equal(m.distance.view(np.uint32), dd.view(np.uint32))
    assert np.array_equal(m.knn_idx, ref.knn_idx) and np.array_equal(m.knn_dist.view(np.uint32), ref.knn_dist.view(np.uint32))
    # a set can be its own partner (self-matching: nearest neighbour of every row is the row itself at distance 0)
    if nq > 1:
        s = ctx.match_two_image(pa, pa, 0.3, want_knn=True)
        assert np.array_equal(s.knn_idx[:, 0], np.arange(nq)) and np.all(s.knn_dist[:, 0] == 0)
    pa.close(); pb.close()
