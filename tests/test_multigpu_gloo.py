"""CPU tests (gloo, world_size 2) of the host-side logic behind the multi-GPU paths: balanced sharding
with no overlap, and the residual-sharded BA protocol -- per-rank normal-equation blocks summed by one
all-reduce equal the single-rank blocks, so every rank takes the same LM step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from spherical_bundle_adjuster_b200 import sharding, synth


def test_shard_ranges_cover_everything_once():
    for n in (0, 1, 7, 64, 2016):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    pairs = sharding.all_pairs(64)
    assert len(pairs) == 2016
    got = sum((sharding.shard_pairs(64, k, 8) for k in range(8)), [])
    assert got == pairs


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_cam, n = 5, 4000
    b1, b2, cam, r_true = synth.make_bearings(n, noise=1e-3, outlier_frac=0.1, seed=3, n_cam=n_cam)
    r = r_true + 0.03
    lo, hi = sharding.shard_range(n, rank, world)
    # the local evaluation of this rank's residual shard (the GPU kernel's job; the oracle stands in on CPU)
    _, _, H, g, cost = oracle.ba_rot_eval(b1[lo:hi], b2[lo:hi], cam[lo:hi], r)
    blk = torch.from_numpy(sharding.pack_blocks(H, g, cost).copy())
    sharding.allreduce_blocks_(blk)
    _, _, Hf, gf, cf = oracle.ba_rot_eval(b1, b2, cam, r)
    full = sharding.pack_blocks(Hf, gf, cf)
    err = float(np.abs(blk.numpy() - full).max() / np.abs(full).max())
    # identical buffers on every rank -> identical (replicated) LM decisions
    gathered = [torch.zeros_like(blk) for _ in range(world)]
    dist.all_gather(gathered, blk)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    # pair sharding: the union over ranks is the whole pair list, in order
    mine = sharding.shard_pairs(10, rank, world)
    counts = [None] * world
    dist.all_gather_object(counts, mine)
    ok_pairs = sum(counts, []) == sharding.all_pairs(10)
    if rank == 0:
        out.put((err, same, ok_pairs))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_residual_sharded_blocks_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(k, 2, port, q)) for k in range(2)]
    for p in procs:
        p.start()
    err, same, ok_pairs = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert err < 1e-12 and same and ok_pairs
