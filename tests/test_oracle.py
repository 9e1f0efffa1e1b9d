"""CPU tests: the oracle (oracle/sba_oracle.c) against the golden vectors in tests/golden/ that were
produced by cv2.BFMatcher, the reference's own equi2cube sources (oracle/_ref) and scipy."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["matcher_64.npz", "matcher_128.npz", "matcher_ragged.npz"])
def test_matcher_oracle_matches_cv2_golden(golden_dir, name):
    g = _load(golden_dir, name)
    idx, dist = oracle.knn2_l2(g["q"], g["t"])
    assert np.array_equal(idx, g["knn_idx"])                                   # bit-exact indices
    assert np.array_equal(dist.view(np.uint32), g["knn_dist"].view(np.uint32))  # bit-exact fp32 distances
    qi, ti, dd = oracle.match_two_image(g["q"], g["t"], 0.3)
    assert np.array_equal(qi, g["keep"])
    assert np.array_equal(ti, g["knn_idx"][g["keep"], 0])


def test_matcher_oracle_matches_live_cv2_when_present():
    cv2 = pytest.importorskip("cv2")
    A, B, _ = synth.make_descriptors(300, 257, 64, seed=5)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(A, B, 2)
    idx, dist = oracle.knn2_l2(A, B)
    assert np.array_equal(idx, np.array([[m.trainIdx for m in k] for k in knn], np.int32))
    assert np.array_equal(dist.view(np.uint32), np.array([[m.distance for m in k] for k in knn], np.float32).view(np.uint32))


def test_matcher_oracle_edge_cases():
    q = synth.unit_rows(np.random.default_rng(0).standard_normal((5, 64))).astype(np.float32)
    idx, dist = oracle.knn2_l2(q, np.zeros((0, 64), np.float32))   # empty train set
    assert (idx == -1).all() and np.isinf(dist).all()
    idx, dist = oracle.knn2_l2(q, q[:1])                            # a single train row: no second neighbour
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()
    assert len(oracle.match_two_image(q, q[:1])[0]) == 0           # guarded: the reference would read knn[i][1]
    idx, _ = oracle.knn2_l2(np.zeros((0, 64), np.float32), q)       # empty query set
    assert idx.shape == (0, 2)


@pytest.mark.parametrize("name", ["remap_even.npz", "remap_odd.npz"])
def test_remap_oracle_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    cs = int(g["cs"])
    assert np.array_equal(oracle.equi2cube_all(g["im"], cs), g["strip"])
    for f in range(6):
        assert np.array_equal(oracle.equi2cube_face(g["im"], cs, f), g["faces"][f])
    # get_all == hconcat(left, front, right, back, top, bottom)  (equi2cube.cpp:293-298)
    assert np.array_equal(np.concatenate(list(g["faces"]), axis=1), g["strip"])


@pytest.mark.parametrize("geom", ["512x256_cs128", "100x50_cs15", "2048x1024_cs512"])
def test_remap_lut_hash(golden_dir, geom):
    h = json.load(open(os.path.join(golden_dir, "remap_lut_hashes.json")))[geom]
    wh, cs = geom.split("_cs")
    w, hh = map(int, wh.split("x"))
    lut, _ = oracle.equi2cube_lut(int(cs), w, hh)
    assert int(lut.astype(np.int64).sum()) == h["sum"]
    assert hashlib.sha256(lut.tobytes()).hexdigest() == h["sha256"]


def test_remap_bottom_centre_clamp():
    # h a power of two and an even cube: theta rounds to pi at the bottom-face centre, the reference
    # indexes row h (equi2cube.cpp:268-275); the restatement clamps that single pixel.
    _, nclamp = oracle.equi2cube_lut(512, 2048, 1024)
    assert nclamp == 1
    _, nclamp = oracle.equi2cube_lut(960, 3840, 1920)
    assert nclamp == 0


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_oracle_equals_compiled_reference():
    im = synth.make_erp_image(256, 128, seed=9)
    for cs in (32, 33, 64):
        assert np.array_equal(oracle.equi2cube_all(im, cs), oracle.ref_equi2cube_all(im, cs))
    rng = np.random.default_rng(3)
    xy = np.stack([rng.uniform(0, 6 * 64, 1000), rng.uniform(0, 64, 1000)], 1).astype(np.float32)
    a, b = oracle.cube2equi_points(xy, 64, 256, 128), oracle.ref_cube2equi_points(xy, 64, 256, 128)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_cube2equi_golden(golden_dir):
    g = _load(golden_dir, "cube2equi.npz")
    out = oracle.cube2equi_points(g["xy"], int(g["cs"]), int(g["w"]), int(g["h"]))
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))


def test_cube2equi_inverts_remap_sampling():
    # property: the ERP pixel of a face-pixel coordinate is the pixel equi2cube sampled
    cs, w, h = 64, 512, 256
    lut, _ = oracle.equi2cube_lut(cs, w, h)
    ii, jj = np.mgrid[0:cs, 0:6 * cs]
    xy = np.stack([jj.ravel() + 0.25, ii.ravel() + 0.25], 1).astype(np.float32)   # strictly inside the pixel
    e = oracle.cube2equi_points(np.stack([jj.ravel(), ii.ravel()], 1).astype(np.float32), cs, w, h)
    src = (np.minimum(e[:, 1].astype(np.int64), h - 1)) * w + np.minimum(e[:, 0].astype(np.int64), w - 1)
    assert (src == lut.ravel()).mean() > 0.995   # float32 rounding of the keypoint flips a few boundary pixels
    assert xy.shape[0] == lut.size


def test_bearings_match_reference_convention():
    # theta from +z, phi from +x (spherical_bundle_adjuster.cpp:279-297)
    xy = np.array([[0, 128], [128, 128], [256, 128], [0, 0], [17.5, 255.9]], np.float32)
    b = oracle.pixels_to_bearings(xy, 512, 256)
    assert np.allclose(b[0], [1, 0, 0], atol=1e-12) and np.allclose(b[1], [0, 1, 0], atol=1e-12)
    assert np.allclose(b[2], [-1, 0, 0], atol=1e-12) and np.allclose(b[3], [0, 0, 1], atol=1e-12)
    assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-14)


def test_ba_oracle_matches_scipy_golden(golden_dir):
    g = _load(golden_dir, "ba_small.npz")
    res, jac, H, gg, cost = oracle.ba_rot_eval(g["b1"], g["b2"], None, g["r"][None], g["t"], float(g["d1"]), float(g["d2"]), 1.0)
    assert np.abs(res - g["res"]).max() < 1e-14
    assert np.abs(jac - g["jac"]).max() < 5e-9            # golden Jacobian is a central difference
    assert np.allclose(H[0], g["H"], rtol=1e-7) and np.allclose(gg[0], g["g"], rtol=1e-6, atol=1e-8)
    assert abs(cost[0] - float(g["cost"])) < 1e-12
    r, s = oracle.ba_rot_solve(g["b1"], g["b2"], None, np.zeros((1, 3)))
    assert np.abs(r - g["r_solved"]).max() < 1e-12 and s.iterations == int(g["iterations"])


def test_ba_oracle_known_answers():
    b1, b2, cam, r_true = synth.make_bearings(500, noise=0.0, seed=1)
    res, *_ = oracle.ba_rot_eval(b1, b2, None, r_true)
    assert np.abs(res).max() < 1e-14                      # zero residual at the truth
    r, s = oracle.ba_rot_solve(b1, b2, None, np.zeros((1, 3)))
    assert np.abs(r - r_true).max() < 1e-6                # recovered rotation <= 1e-6 rad
    # small-angle branch (theta^2 <= DBL_EPSILON): out = p + r x p
    rs = np.array([[1e-9, -2e-9, 3e-9]])
    res_s, jac_s, *_ = oracle.ba_rot_eval(b1, b1, None, rs)
    assert np.allclose(res_s, -np.cross(rs[0], b1), atol=1e-20)
    # theta near pi
    rp = np.array([[np.pi - 1e-7, 0.0, 0.0]])
    res_p, *_ = oracle.ba_rot_eval(b1, b1 * [1, -1, -1], None, rp)
    assert np.abs(res_p).max() < 1e-6
    # Huber: outliers beyond |res| > 1 contribute 2|res|-1
    b2o = -b1
    _, _, _, _, cost = oracle.ba_rot_eval(b1, b2o, None, np.zeros((1, 3)))
    assert abs(cost[0] - 0.5 * 500 * (2 * 2.0 - 1)) < 1e-9


def test_ba_oracle_multi_camera_blocks_are_independent():
    b1, b2, cam, r_true = synth.make_bearings(3000, noise=1e-3, seed=3, n_cam=4)
    r0 = np.zeros((4, 3))
    _, _, H, g, cost = oracle.ba_rot_eval(b1, b2, cam, r0)
    for c in range(4):
        m = cam == c
        _, _, Hc, gc, cc = oracle.ba_rot_eval(b1[m], b2[m], None, r0[c:c + 1])
        assert np.allclose(H[c], Hc[0], rtol=1e-12) and np.allclose(g[c], gc[0], rtol=1e-10, atol=1e-12)
    r, s = oracle.ba_rot_solve(b1, b2, cam, r0)
    assert np.abs(r - r_true).max() < 5e-4


def test_ba_tran_only_oracle():
    # translation block (spherical_bundle_adjuster.cpp:948-1002): d res / d t = +I
    b1, b2, cam, r_true = synth.make_bearings(800, noise=0.0, seed=2)
    t_true = np.array([[0.03, -0.02, 0.01]])
    X2 = b2 - t_true
    res, H, g, cost = oracle.ba_tran_eval(b1, X2, None, r_true, t_true)
    assert np.abs(res).max() < 1e-14 and np.allclose(H[0], [800, 0, 0, 800, 0, 800])
    tv, s = oracle.ba_tran_solve(b1, X2, None, r_true, np.zeros((1, 3)))
    assert np.abs(tv - t_true).max() < 1e-9
    eps = 1e-6
    t0 = np.array([0.1, 0.2, -0.1])
    c = lambda t: oracle.ba_tran_eval(b1, b2, None, r_true, t[None])[3][0]
    fd = np.array([(c(t0 + eps * np.eye(3)[k]) - c(t0 - eps * np.eye(3)[k])) / (2 * eps) for k in range(3)])
    assert np.abs(fd - oracle.ba_tran_eval(b1, b2, None, r_true, t0[None])[2][0]).max() < 1e-6


def test_ba_depth_only_functor_oracle():
    # depth block (spherical_bundle_adjuster.cpp:1005-1032): Jacobian against central differences
    b1, b2, r, t, _ = synth.make_two_view(20, seed=1)
    d = np.array([1.3, 0.4])
    for i in range(20):
        res, J = oracle.ba_d_functor(b1[i], b2[i], r, t, d, 1.5, 0.8)
        assert res.shape == (5,) and abs(res[3] - 1.5 * np.exp(-0.8 * 1.3)) < 1e-15
        for k in range(2):
            e = np.zeros(2); e[k] = 1e-6
            fd = (oracle.ba_d_functor(b1[i], b2[i], r, t, d + e, 1.5, 0.8)[0] - oracle.ba_d_functor(b1[i], b2[i], r, t, d - e, 1.5, 0.8)[0]) / 2e-6
            assert np.abs(fd - J[:, k]).max() < 1e-9


def _hermite_min(samples, lo, hi):
    """Independent restatement with numpy: fit value+gradient samples, minimise over [lo, hi] the way
    Ceres' MinimizePolynomial does (mid point, ends, real parts of the derivative's roots)."""
    nc = 2 * len(samples)
    A, b = [], []
    for x, v, g in samples:
        A.append([x ** (nc - 1 - j) for j in range(nc)]); b.append(v)
        A.append([(nc - 1 - j) * x ** (nc - 2 - j) if j < nc - 1 else 0.0 for j in range(nc)]); b.append(g)
    p = np.linalg.solve(np.array(A), np.array(b))
    cands = [0.5 * (lo + hi), lo, hi] + [z.real for z in np.roots(np.polyder(p)) if lo <= z.real <= hi]
    vals = [np.polyval(p, x) for x in cands]
    return cands[int(np.argmin(vals))]


def test_line_search_interpolation_oracle():
    # two-sample (cubic) and three-sample (quintic) fits, incl. a case whose derivative has complex roots
    cases = [
        ((10.0, -4.0), (1.0, 12.0, 9.0), None),
        ((10.0, -4.0), (0.3, 9.9, 2.5), (1.0, 12.0, 9.0)),
        ((5.0, -1.0), (1.0, 5.5, 0.2), None),
        ((5.0, -1.0), (0.6, 5.2, -0.1), (1.0, 5.5, 0.2)),
        ((1.0, -100.0), (1.0, 50.0, 400.0), None),
    ]
    for (f0, g0), cur, prev in cases:
        got = oracle.ls_next_step(f0, g0, cur, prev)
        samples = [(0.0, f0, g0), cur] + ([prev] if prev is not None else [])
        want = _hermite_min(samples, 1e-3 * cur[0], 0.6 * cur[0])
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (got, want)


def test_ba_depth_only_solve_oracle():
    from scipy.optimize import least_squares
    b1, b2, r, t, depths = synth.make_two_view(60, seed=3, outlier_frac=0.1)
    d, s, nls = oracle.ba_d_solve(b1, b2, r, t, np.full((60, 2), 1.0))
    assert s.termination == 1 and s.final_cost < s.initial_cost and np.all(d >= 0)
    # every block is independent: scipy's bounded trust-region solver on single blocks lands at the same
    # minimiser up to the function-tolerance slack Ceres stops with
    for i in range(0, 60, 7):
        sol = least_squares(lambda x: oracle.ba_d_functor(b1[i], b2[i], r, t, x)[0], [1.0, 1.0], bounds=(0, np.inf), xtol=1e-14, ftol=1e-14, gtol=1e-14)
        assert np.abs(sol.x - d[i]).max() < 2e-2
    # no barrier: outliers are pushed onto the bound and stay feasible
    d2, s2, _ = oracle.ba_d_solve(b1, b2, r, t, np.full((60, 2), 1.0), 0.0, 1.0)
    assert np.all(d2 >= 0) and (d2 == 0).any()
    # infeasible start is projected first (IterationZero)
    d3, s3, _ = oracle.ba_d_solve(b1, b2, r, t, np.full((60, 2), -2.0))
    d4, s4, _ = oracle.ba_d_solve(b1, b2, r, t, np.zeros((60, 2)))
    assert np.array_equal(d3, d4) and s3.iterations == s4.iterations
    # stiff barrier exercises the Armijo interpolation
    b1, b2, r, t, _ = synth.make_two_view(400, seed=400)
    d5, s5, nls5 = oracle.ba_d_solve(b1, b2, r, t, np.full((400, 2), 5.0), 20.0, 8.0)
    assert nls5 > 0 and s5.final_cost < s5.initial_cost
    # iteration cap
    _, s6, _ = oracle.ba_d_solve(b1, b2, r, t, np.full((400, 2), 5.0), 1.0, 1.0, 3)
    assert s6.iterations == 3 and s6.termination == 0


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_spherical_surf_oracle_pinned_to_reference():
    # spherical_surf.cpp:17-123 restated in oracle/sba_oracle.c vs the reference's own file compiled into oracle/_ref
    for th in ([0, 0.3, 0], [0.1, -0.7, 1.2], [0, np.float32(np.pi * 45 / 180), 0]):
        assert np.array_equal(oracle.eular2rot(th), oracle.ref_eular2rot(th))
    rng = np.random.default_rng(0)
    for (w, h) in ((1024, 512), (1000, 500), (258, 130)):
        im = synth.make_erp_image(w, h, seed=1)
        rc = np.stack([rng.integers(0, h, 500), rng.integers(0, w, 500)], axis=1).astype(np.int32)
        xy = (rng.uniform(0, 1, (800, 2)) * [w - 1, h / 4 - 1]).astype(np.float32)
        for pitch in (45.0, -45.0, -90.0, 30.5):
            assert np.array_equal(oracle.crop_rotated_image(im, pitch), oracle.ref_crop_rotated_image(im, pitch))
            assert np.array_equal(oracle.rotate_pixels(rc, pitch, w, h), oracle.ref_rotate_pixels(rc, pitch, w, h))
            assert np.array_equal(oracle.rotate_keypoints(xy, pitch, w, h), oracle.ref_rotate_keypoints(xy, pitch, w, h))
    # the -90 degree crop of a 1024x512 image has a pixel whose source is undefined (NaN): skipped by the reference
    assert (oracle.crop_rotated_lut(-90.0, 1024, 512) < 0).sum() >= 1


def test_spherical_surf_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "spherical_surf.npz"))
    w, h = int(g["w"]), int(g["h"])
    im = synth.make_erp_image(w, h, seed=int(g["seed"]))
    for k, pitch in enumerate(g["pitches"]):
        lut = oracle.crop_rotated_lut(float(pitch), w, h)
        assert hashlib.sha256(lut.tobytes()).hexdigest() == str(g["lut_sha256"][k])
        assert np.array_equal(oracle.crop_rotated_image(im, float(pitch))[::37, ::41], g["crop_samples"][k])
        assert np.array_equal(oracle.rotate_keypoints(g["keys"], float(pitch), w, h), g["keys_rotated"][k])


def test_initial_guess_oracle_pinned_to_opencv():
    """eight_point_estimation's linear algebra (spherical_bundle_adjuster.cpp:53-82) against the installed OpenCV:
    cv2.SVDecomp and cv2.decomposeEssentialMat are the calls the reference makes."""
    cv2 = pytest.importorskip("cv2")
    b1, b2, r, t, _ = synth.make_two_view(2000, seed=3, outlier_frac=0.05)
    rng = np.random.default_rng(0)
    idx = np.stack([rng.permutation(2000)[:500] for _ in range(12)]).astype(np.int32)

    def rot2euler(R):
        sy = np.float32(np.sqrt(R[0, 0] ** 2 + R[1, 0] ** 2))
        return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arctan2(-R[2, 0], sy), np.arctan2(R[1, 0], R[0, 0])], np.float32)

    for s in range(len(idx)):
        A = np.einsum("na,nb->nab", b1[idx[s]], b2[idx[s]]).reshape(-1, 9)
        w, u, vt = cv2.SVDecomp(A)
        e, sv = oracle.eight_point_null(b1, b2, idx[s])
        assert np.abs(e * np.sign(e @ vt[-1]) - vt[-1]).max() < 1e-12 and np.allclose(sv, w.ravel(), rtol=1e-10)
        w2, u2, vt2 = cv2.SVDecomp(vt[-1].reshape(3, 3))
        w2[2] = 0
        R1, R2, tt = cv2.decomposeEssentialMat(u2 @ np.diag(w2.ravel()) @ vt2)
        o1, o2, oT, v1, v2 = oracle.essential_to_candidates(e)
        c1, c2 = rot2euler(R1), rot2euler(R2)
        same = max(np.abs(o1 - c1).max(), np.abs(o2 - c2).max())
        swap = max(np.abs(o1 - c2).max(), np.abs(o2 - c1).max())
        assert min(same, swap) < 1e-6 and np.abs(np.abs(oT) - np.abs(tt.ravel())).max() < 1e-6
    # the vote: hand-made candidates, the tight cluster wins over the stragglers
    cand = np.concatenate([np.float32([0.1, 0.2, 0.3]) + 1e-3 * rng.standard_normal((30, 3)).astype(np.float32),
                           rng.uniform(-1.5, 1.5, (10, 3)).astype(np.float32)])
    win = oracle.lib().orc_vote_rotation(oracle._p(np.ascontiguousarray(cand), __import__("ctypes").c_float), len(cand))
    assert win < 30
    R, T, best, c = oracle.initial_guess(b1, b2, idx)
    assert best >= 0 and len(c) >= len(idx)


def test_spherical_surf_geometry_properties():
    """Size-independent properties of the pitched-band mapping (spherical_surf.cpp:48-109)."""
    rng = np.random.default_rng(7)
    for (w, h) in ((512, 256), (1000, 500), (3840, 1920)):
        for pitch in (45.0, -45.0, -90.0, 17.5):
            lut = oracle.crop_rotated_lut(pitch, w, h)
            assert lut.shape == (h // 4, w) and lut.max() < w * h and lut.min() >= -1
            # rotating by the pitch and back lands within a pixel (two truncations) of where it started, modulo the
            # azimuth wrap -- away from the poles, where a column step is a large angle
            rc = np.stack([rng.integers(h * 3 // 8, h * 5 // 8, 300), rng.integers(0, w, 300)], axis=1).astype(np.int32)
            fwd = oracle.rotate_pixels(rc, pitch, w, h)
            ok = (fwd[:, 0] > h // 8) & (fwd[:, 0] < h * 7 // 8)
            back = oracle.rotate_pixels(fwd[ok], -pitch, w, h)
            dr = np.abs(back[:, 0] - rc[ok, 0])
            dc = np.abs(back[:, 1] - rc[ok, 1]); dc = np.minimum(dc, w - dc)
            assert dr.max() <= 2 and dc.max() <= 3, (w, h, pitch, dr.max(), dc.max())
        # pitch 0 is the identity on the band up to the truncation of the round trip through radians (which is why the
        # reference takes the plain ROI for its unrotated band instead, spherical_surf.cpp:139)
        ident = oracle.crop_rotated_lut(0.0, w, h)
        rows, cols = ident // w, ident % w
        want_r = (np.arange(h // 4)[:, None] + h * 3 // 8) * np.ones((1, w), np.int64)
        want_c = np.ones((h // 4, 1), np.int64) * np.arange(w)[None, :]
        assert np.abs(rows - want_r).max() <= 1 and np.abs(cols - want_c).max() <= 1
        assert (rows == want_r).mean() > 0.9 and (cols == want_c).mean() > 0.75


def test_depth_block_is_separable():
    """Every match is its own block: solving a subset alone reaches the same depths as inside the full problem when
    both runs are driven to the same number of (all accepted) iterations is not guaranteed -- but the fully converged
    optimum is, so compare optima reached with generous iteration caps."""
    b1, b2, r, t, _ = synth.make_two_view(40, seed=9, outlier_frac=0.0, noise=1e-3)
    d_all, s_all, _ = oracle.ba_d_solve(b1, b2, r, t, np.full((40, 2), 1.0), 1.0, 1.0, 500)
    d_sub, s_sub, _ = oracle.ba_d_solve(b1[:10], b2[:10], r, t, np.full((10, 2), 1.0), 1.0, 1.0, 500)
    assert np.abs(d_all[:10] - d_sub).max() < 5e-2 * np.abs(d_sub).max()
    # cost is the sum of the per-match costs
    per = sum(0.5 * (oracle.ba_d_functor(b1[i], b2[i], r, t, d_all[i])[0] ** 2).sum() for i in range(40))
    assert abs(per - s_all.final_cost) < 1e-9 * max(1.0, per)
