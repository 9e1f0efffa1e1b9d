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


@pytest.mark.parametrize("kind", ["host", "device"])
def test_pair_rotation_single_call(ctx, kind):
    """sba_pair_rotation (one C-ABI call: remap x2 + match + cube2equi + bearings + rotation BA) against
    the oracle run stage by stage, with host buffers and with device tensors."""
    import torch
    w, h, cs, n = 1024, 512, 256, 3000
    pair = synth.make_pair(n, n - 77, cs=cs, seed=12, rotvec=(-0.2, 0.15, 0.4))
    im1, im2 = synth.make_erp_image(w, h, seed=2), synth.make_erp_image(w, h, seed=3)
    args = [im1, im2, pair["desc1"], pair["desc2"], pair["key1_xy"], pair["key2_xy"]]
    if kind == "device":
        args = [torch.from_numpy(a).cuda() for a in args]
    res, (qi, ti, dd), (s1, s2) = ctx.pair_rotation(*args, cs, want_strips=True)
    if kind == "device":
        torch.cuda.synchronize()
        qi, ti, dd, s1, s2 = [x.cpu().numpy() for x in (qi, ti, dd, s1, s2)]
    assert np.array_equal(s1, oracle.equi2cube_all(im1, cs)) and np.array_equal(s2, oracle.equi2cube_all(im2, cs))
    oqi, oti, odd = oracle.match_two_image(pair["desc1"], pair["desc2"], 0.3)
    assert res.n_matches == len(oqi) and np.array_equal(qi, oqi) and np.array_equal(ti, oti)
    assert np.array_equal(dd.view(np.uint32), odd.view(np.uint32))
    e1 = oracle.cube2equi_points(pair["key1_xy"][oqi], cs, w, h); e2 = oracle.cube2equi_points(pair["key2_xy"][oti], cs, w, h)
    b1 = oracle.pixels_to_bearings(e1, w, h).astype(np.float32).astype(np.float64)
    b2 = oracle.pixels_to_bearings(e2, w, h).astype(np.float32).astype(np.float64)
    r_or, s_or = oracle.ba_rot_solve(b1, b2, None, np.zeros((1, 3)))
    r = np.array(res.rotation)
    assert np.abs(r - r_or[0]).max() < 1e-6 and res.lm_iterations == s_or.iterations
    assert np.linalg.norm(r - pair["r_true"]) < 1e-4


def test_pair_rotation_no_matches_and_no_images(ctx):
    rng = np.random.default_rng(0)
    q = synth.unit_rows(rng.standard_normal((200, 64))).astype(np.float32)
    t = synth.unit_rows(rng.standard_normal((300, 64))).astype(np.float32)      # unrelated sets: nothing survives
    k1 = rng.uniform(0, 100, (200, 2)).astype(np.float32); k2 = rng.uniform(0, 100, (300, 2)).astype(np.float32)
    res, (qi, ti, dd), strips = ctx.pair_rotation(None, None, q, t, k1, k2, 128, w=512, h=256, r0=(0.1, 0.2, 0.3))
    assert res.n_matches == 0 and len(qi) == 0 and strips is None
    assert np.allclose(res.rotation, (0.1, 0.2, 0.3))                            # initial value returned untouched


def test_pair_rotation_repeatable_across_buffer_growth(ctx):
    """The same pair gives bit-identical results before and after a LARGER call on the same context forced every
    scratch buffer to be reallocated (the hazard the removed graph-replay path had: stale addresses)."""
    w, h, cs = 1024, 512, 256
    def run(n, seed):
        pair = synth.make_pair(n, n, cs=cs, seed=seed, rotvec=(0.1, -0.2, 0.3))
        im1, im2 = synth.make_erp_image(w, h, seed=2), synth.make_erp_image(w, h, seed=3)
        res, (qi, ti, dd), _ = ctx.pair_rotation(im1, im2, pair["desc1"], pair["desc2"], pair["key1_xy"], pair["key2_xy"], cs)
        return tuple(res.rotation), res.n_matches, res.lm_iterations, qi.copy(), ti.copy(), dd.copy()
    a = run(2000, 5)
    run(9000, 6)                                        # grows SCR_PIPE_*, the matcher workspace and the BA blocks
    b = run(2000, 5)
    assert a[:3] == b[:3] and a[1] > 500
    for x, y in zip(a[3:], b[3:]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("kind", ["host", "device"])
def test_pair_rotation_single_call(ctx, kind):
    """sba_pair_rotation (one C-ABI call: remap x2 + match + cube2equi + bearings + rotation BA) against
    the oracle run stage by stage, with host buffers and with device tensors."""
    import torch
    w, h, cs, n = 1024, 512, 256, 3000
    pair = synth.make_pair(n, n - 77, cs=cs, seed=12, rotvec=(-0.2, 0.15, 0.4))
    im1, im2 = synth.make_erp_image(w, h, seed=2), synth.make_erp_image(w, h, seed=3)
    args = [im1, im2, pair["desc1"], pair["desc2"], pair["key1_xy"], pair["key2_xy"]]
    if kind == "device":
        args = [torch.from_numpy(a).cuda() for a in args]
    res, (qi, ti, dd), (s1, s2) = ctx.pair_rotation(*args, cs, want_strips=True)
    if kind == "device":
        torch.cuda.synchronize()
        qi, ti, dd, s1, s2 = [x.cpu().numpy() for x in (qi, ti, dd, s1, s2)]
    assert np.array_equal(s1, oracle.equi2cube_all(im1, cs)) and np.array_equal(s2, oracle.equi2cube_all(im2, cs))
    oqi, oti, odd = oracle.match_two_image(pair["desc1"], pair["desc2"], 0.3)
    assert res.n_matches == len(oqi) and np.array_equal(qi, oqi) and np.array_equal(ti, oti)
    assert np.array_equal(dd.view(np.uint32), odd.view(np.uint32))
    e1 = oracle.cube2equi_points(pair["key1_xy"][oqi], cs, w, h); e2 = oracle.cube2equi_points(pair["key2_xy"][oti], cs, w, h)
    b1 = oracle.pixels_to_bearings(e1, w, h).astype(np.float32).astype(np.float64)
    b2 = oracle.pixels_to_bearings(e2, w, h).astype(np.float32).astype(np.float64)
    r_or, s_or = oracle.ba_rot_solve(b1, b2, None, np.zeros((1, 3)))
    r = np.array(res.rotation)
    assert np.abs(r - r_or[0]).max() < 1e-6 and res.lm_iterations == s_or.iterations
    assert np.linalg.norm(r - pair["r_true"]) < 1e-4


def test_pair_rotation_no_matches_and_no_images(ctx):
    rng = np.random.default_rng(0)
    q = synth.unit_rows(rng.standard_normal((200, 64))).astype(np.float32)
    t = synth.unit_rows(rng.standard_normal((300, 64))).astype(np.float32)      # unrelated sets: nothing survives
    k1 = rng.uniform(0, 100, (200, 2)).astype(np.float32); k2 = rng.uniform(0, 100, (300, 2)).astype(np.float32)
    res, (qi, ti, dd), strips = ctx.pair_rotation(None, None, q, t, k1, k2, 128, w=512, h=256, r0=(0.1, 0.2, 0.3))
    assert res.n_matches == 0 and len(qi) == 0 and strips is None
    assert np.allclose(res.rotation, (0.1, 0.2, 0.3))                            # initial value returned untouched


def test_pair_rotation_dependent_launch_gives_identical_results():
    """sba_ctx_set_dependent_launch: the pair's kernel chain launched programmatically dependent (each kernel scheduled while
    its predecessor drains, waiting on the device before it touches memory) -- same match list, same distance bits, same
    rotation bits as the plain stream-ordered launches, call after call, on device tensors and on host buffers; the match
    entry point alone as well."""
    import torch
    from spherical_bundle_adjuster_b200 import Context
    c = Context(0)
    w, h, cs, n = 1024, 512, 256, 3000
    pair = synth.make_pair(n, n, cs=cs, seed=5, rotvec=(0.1, -0.2, 0.3))
    im1, im2 = synth.make_erp_image(w, h, seed=2), synth.make_erp_image(w, h, seed=3)
    host = [im1, im2, pair["desc1"], pair["desc2"], pair["key1_xy"], pair["key2_xy"]]
    dev = [torch.from_numpy(a).cuda() for a in host]
    runs = {}
    for flag in (False, True, False, True):
        c.set_dependent_launch(flag)
        out = []
        for k in range(6):
            args = dev if k % 2 == 0 else host
            res, (qi, ti, dd), _ = c.pair_rotation(*args, cs, r0=(0.01 * (k % 3), 0.0, 0.0), want_matches=True)
            if k % 2 == 0:
                torch.cuda.synchronize()
                qi, ti, dd = qi.cpu().numpy(), ti.cpu().numpy(), dd.cpu().numpy()
            out.append((tuple(res.rotation), res.n_matches, res.lm_iterations, res.lm_termination, qi.tobytes(), ti.tobytes(), dd.tobytes()))
        m = c.match_two_image(dev[2], dev[3], 0.3)
        torch.cuda.synchronize()
        out.append(tuple(x.cpu().numpy().tobytes() for x in (m.query_idx, m.train_idx, m.distance)))
        runs.setdefault(flag, []).append(out)
    c.set_dependent_launch(False)
    assert runs[False][0] == runs[False][1] == runs[True][0] == runs[True][1]
    assert runs[False][0][0][1] > 1000


@pytest.mark.parametrize("kind", ["host", "device"])
def test_pair_rotation_begin_end_keeps_pairs_in_flight(kind):
    """sba_pair_rotation_begin/_end: three pairs queued on three contexts (own streams) before any result is
    collected give exactly what the blocking call gives; a second begin on a busy context is refused."""
    import torch
    from spherical_bundle_adjuster_b200 import Context, SbaError
    w, h, cs, n = 1024, 512, 256, 2500
    streams = [torch.cuda.Stream() for _ in range(3)]
    ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
    data = []
    for k in range(3):
        pair = synth.make_pair(n, n - 31 * k, cs=cs, seed=40 + k, rotvec=(0.1 * k - 0.1, 0.2, -0.3))
        a = [synth.make_erp_image(w, h, seed=2 * k), synth.make_erp_image(w, h, seed=2 * k + 1), pair["desc1"], pair["desc2"], pair["key1_xy"],
             pair["key2_xy"]]
        if kind == "device":
            a = [torch.from_numpy(x).cuda() for x in a]
        else:
            a = [torch.from_numpy(x).pin_memory() for x in a]
        data.append(a)
    torch.cuda.synchronize()
    want = [ctxs[0].pair_rotation(*a, cs, want_strips=True) for a in data]
    calls = [ctxs[k].pair_rotation_begin(*data[k], cs, want_strips=True) for k in range(3)]
    with pytest.raises(SbaError):
        ctxs[1].pair_rotation_begin(*data[0], cs)
    got = [c.end() for c in reversed(calls)][::-1]          # collect in another order than queued
    torch.cuda.synchronize()
    tonp = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    for (r0, m0, s0), (r1, m1, s1) in zip(want, got):
        assert r0.n_matches == r1.n_matches > 0 and np.array_equal(np.array(r0.rotation), np.array(r1.rotation))
        for a, b in zip(m0 + s0, m1 + s1):
            assert np.array_equal(tonp(a), tonp(b))
    again = ctxs[1].pair_rotation_begin(*data[0], cs).end()   # the context is free again after end()
    assert again[0].n_matches == want[0][0].n_matches
    for c in ctxs:
        c.close()
