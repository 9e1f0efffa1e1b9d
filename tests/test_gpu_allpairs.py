"""GPU parity for BASELINE config 3 (all-pairs matching of a frame sequence, pairs sharded across ranks): every pair of
every shard, matched with two library contexts in flight, equals the oracle bit for bit."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import Context, sharding, synth

pytestmark = pytest.mark.gpu


def test_all_pairs_sharded_matches_oracle():
    import torch
    F, n = 5, 1100                                   # 10 pairs; 1100 x 1100 takes the tensor-core path
    rng = np.random.default_rng(3)
    base = synth.unit_rows(rng.standard_normal((n, 64)))
    frames = [synth.unit_rows(base[rng.permutation(n)] + 0.02 * rng.standard_normal((n, 64))).astype(np.float32) for _ in range(F)]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    streams = [torch.cuda.Stream() for _ in range(2)]
    ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
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
