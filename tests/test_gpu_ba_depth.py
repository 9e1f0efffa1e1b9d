"""GPU parity: depth-only block (spherical_bundle_adjuster.cpp:1005-1063) and the three-stage solve_problem
(:183-217) through the C ABI against the fp64 oracle.

Tolerances: functor values 1e-12 relative (same fp64 arithmetic up to summation order and libm);
solved depths 1e-7 * max(1, |d|); the solver's discrete trajectory (iterations, accepted steps,
termination reason, line-search trials) must be identical."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu


def _problem(ctx, n, seed, outliers=0.05):
    b1, b2, r, t, depths = synth.make_two_view(n, outlier_frac=outliers, seed=seed)
    b1f, b2f = b1.astype(np.float32), b2.astype(np.float32)
    return ctx.ba_problem(b1f, b2f), b1f.astype(np.float64), b2f.astype(np.float64), r, t, depths


@pytest.mark.parametrize("n", [1, 33, 5000])
def test_depth_functor(ctx, n):
    prob, b1, b2, r, t, depths = _problem(ctx, n, seed=n)
    d = np.abs(np.random.default_rng(n).normal(1.0, 0.5, (n, 2)))
    res, jac, cost = prob.d_eval(r, t, d, 1.3, 0.7)
    for i in range(0, n, max(1, n // 50)):
        rr, J = oracle.ba_d_functor(b1[i], b2[i], r, t, d[i], 1.3, 0.7)
        assert np.allclose(res[i], rr, rtol=1e-12, atol=1e-13) and np.allclose(jac[i], J, rtol=1e-12, atol=1e-13)
        assert abs(cost[i] - 0.5 * rr @ rr) <= 1e-12 * max(1.0, cost[i])


CASES = [
    # n, start depth, lambda, c, outlier share
    (2, 1.0, 1.0, 1.0, 0.0),
    (50, 1.0, 1.0, 1.0, 0.05),          # the reference's constants
    (400, 5.0, 1.0, 1.0, 0.05),
    (3000, 0.0, 1.0, 1.0, 0.05),        # starts ON the bound
    (3000, -3.0, 1.0, 1.0, 0.05),       # infeasible start: projected first
    (400, 1.0, 0.0, 1.0, 0.3),          # no barrier: outliers end on the bound d = 0
    (400, 5.0, 20.0, 8.0, 0.05),        # stiff barrier: the Armijo search has to shorten steps
    (3000, 1.0, 20.0, 8.0, 0.05),
    (100000, 1.0, 1.0, 1.0, 0.1),
]


@pytest.mark.parametrize("n,d0,lam,c,outliers", CASES)
def test_depth_solve_matches_oracle(ctx, n, d0, lam, c, outliers):
    prob, b1, b2, r, t, _ = _problem(ctx, n, seed=n, outliers=outliers)
    start = np.full((n, 2), d0)
    d_ref, s_ref, nls_ref = oracle.ba_d_solve(b1, b2, r, t, start, lam, c)
    d, s, nls = prob.d_solve(r, t, start, lam, c)
    assert (s.iterations, s.num_successful, s.termination, nls) == (s_ref.iterations, s_ref.num_successful, s_ref.termination, nls_ref)
    assert abs(s.initial_cost - s_ref.initial_cost) <= 1e-11 * max(1.0, s_ref.initial_cost)
    assert abs(s.final_cost - s_ref.final_cost) <= 1e-9 * max(1.0, s_ref.final_cost)
    assert np.all(d >= 0.0)
    assert np.all(np.abs(d - d_ref) <= 1e-7 * np.maximum(1.0, np.abs(d_ref)))
    assert np.array_equal(d == 0.0, d_ref == 0.0)


def test_depth_solve_line_search_ran(ctx):
    """The stiff-barrier case must actually exercise the interpolation code on the device."""
    prob, b1, b2, r, t, _ = _problem(ctx, 400, seed=400)
    _, s, nls = prob.d_solve(r, t, np.full((400, 2), 5.0), 20.0, 8.0)
    assert nls > 0 and s.evaluations > s.iterations


def test_depth_solve_iteration_cap(ctx):
    prob, b1, b2, r, t, _ = _problem(ctx, 400, seed=7)
    start = np.full((400, 2), 5.0)
    for cap in (0, 1, 3):
        d_ref, s_ref, _ = oracle.ba_d_solve(b1, b2, r, t, start, 1.0, 1.0, cap)
        d, s, _ = prob.d_solve(r, t, start, 1.0, 1.0, cap)
        assert (s.iterations, s.termination) == (s_ref.iterations, s_ref.termination) and s.iterations <= cap
        assert np.all(np.abs(d - d_ref) <= 1e-9 * np.maximum(1.0, np.abs(d_ref)))


def test_depth_solve_recovers_depths(ctx):
    """Noise-free inliers: the solved depths are the true ranges up to the small bias of the barrier terms."""
    b1, b2, r, t, depths = synth.make_two_view(2000, noise=0.0, outlier_frac=0.0, seed=11)
    prob = ctx.ba_problem(b1.astype(np.float32), b2.astype(np.float32))
    d, s, _ = prob.d_solve(r, t, np.full((2000, 2), 1.0), 1.0, 1.0, max_iter=200)
    assert s.termination in (1, 2, 3)
    assert np.median(np.abs(d - depths) / depths) < 0.05


@pytest.mark.parametrize("n", [300, 20000])
def test_solve_problem_three_stages(ctx, n):
    prob, b1, b2, r_true, t_true, _ = _problem(ctx, n, seed=n + 1)
    r0 = r_true + np.array([0.02, -0.03, 0.01])
    t0 = t_true + np.array([0.05, -0.02, 0.03])
    d0 = np.full((n, 2), 1.0)
    # oracle: the same sequence, stage by stage (spherical_bundle_adjuster.cpp:196-209)
    d_ref, s_d, _ = oracle.ba_d_solve(b1, b2, r0, t0, d0)
    r_ref, s_r = oracle.ba_rot_solve(b1, b2, None, r0[None], t0, d_ref[0, 0], d_ref[1, 0], 1.0)
    t_ref, s_t = oracle.ba_tran_solve(b1, b2, None, r_ref, t0[None], d_ref[0, 0], d_ref[1, 0], 1.0)
    r, t, d, sums = prob.solve_problem(r0, t0, d0)
    assert np.all(np.abs(d - d_ref) <= 1e-7 * np.maximum(1.0, np.abs(d_ref)))
    assert np.abs(r - r_ref[0]).max() < 1e-6            # rotations within 1e-6 rad
    assert np.abs(t - t_ref[0]).max() < 1e-6
    assert sums[0].iterations == s_d.iterations and sums[0].termination == s_d.termination
    assert sums[1].termination == s_r.termination and sums[2].termination == s_t.termination


def test_depth_rejects_multi_camera(ctx):
    b1, b2, cam, _ = synth.make_bearings(64, n_cam=2, seed=3)
    prob = ctx.ba_problem(b1.astype(np.float32), b2.astype(np.float32), cam, 2)
    with pytest.raises(Exception):
        prob.d_solve(np.zeros(3), np.zeros(3), np.ones((64, 2)))


def test_depth_and_three_stage_solve_with_device_buffers(ctx):
    """The depth table may live on the device (SBA_MEM_DEVICE): same results as with host buffers."""
    import torch
    n = 4000
    prob, b1, b2, r_true, t_true, _ = _problem(ctx, n, seed=21)
    r0, t0 = r_true + np.array([0.01, 0.02, -0.01]), t_true + np.array([-0.02, 0.01, 0.02])
    d0 = np.full((n, 2), 1.0)
    d_h, s_h, _ = prob.d_solve(r0, t0, d0)
    d_d, s_d, _ = prob.d_solve(r0, t0, torch.from_numpy(d0).cuda())
    assert d_d.is_cuda and np.array_equal(d_d.cpu().numpy(), d_h) and s_d.iterations == s_h.iterations
    rh, th, dh, _ = prob.solve_problem(r0, t0, d0)
    rd, td, dd, _ = prob.solve_problem(r0, t0, torch.from_numpy(d0).cuda())
    assert np.array_equal(rh, rd) and np.array_equal(th, td) and np.array_equal(dd.cpu().numpy(), dh)
