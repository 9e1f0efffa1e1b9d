"""GPU: the whole hot path through the C ABI (remap -> match -> cube2equi -> bearings -> rotation BA)
on a synthetic ERP pair, device-resident, checked stage by stage against the oracle."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_pair_pipeline_device_resident(ctx):
    import torch
    w, h, cs, n = 2048, 1024, 512, 4096
    pair = synth.make_pair(n, n, cs=cs, seed=3)
    im = synth.make_erp_image(w, h, seed=1)
    d = lambda a: torch.from_numpy(a).cuda()
    strip = ctx.equi2cube(d(im), cs)
    m = ctx.match_two_image(d(pair["desc1"]), d(pair["desc2"]), 0.3)
    kl, kr = ctx.gather_matches(d(pair["key1_xy"]), d(pair["key2_xy"]), m.query_idx, m.train_idx)
    b1 = ctx.pixels_to_bearings(ctx.cube2equi_points(kl, cs, w, h), w, h)
    b2 = ctx.pixels_to_bearings(ctx.cube2equi_points(kr, cs, w, h), w, h)
    prob = ctx.ba_problem(b1, b2)
    r, s = prob.solve(np.zeros((1, 3)))
    torch.cuda.synchronize()

    assert np.array_equal(strip.cpu().numpy(), oracle.equi2cube_all(im, cs))
    qi, ti, _ = oracle.match_two_image(pair["desc1"], pair["desc2"], 0.3)
    assert np.array_equal(m.query_idx.cpu().numpy(), qi) and np.array_equal(m.train_idx.cpu().numpy(), ti)
    ob1 = oracle.pixels_to_bearings(oracle.cube2equi_points(pair["key1_xy"][qi], cs, w, h), w, h)
    ob2 = oracle.pixels_to_bearings(oracle.cube2equi_points(pair["key2_xy"][ti], cs, w, h), w, h)
    assert np.abs(b1.cpu().numpy()[:, :3] - ob1).max() < 1e-6
    # the oracle solves on the same fp32-rounded bearings the device problem holds
    r_or, _ = oracle.ba_rot_solve(b1.cpu().numpy()[:, :3].astype(np.float64), b2.cpu().numpy()[:, :3].astype(np.float64), None, np.zeros((1, 3)))
    assert np.abs(r - r_or).max() < 1e-6
    assert np.linalg.norm(r[0] - pair["r_true"]) < 1e-4
