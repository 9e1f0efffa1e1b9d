"""GPU parity: eight_point_estimation / initial_guess (spherical_bundle_adjuster.cpp:47-181) through the C ABI
against the oracle (whose SVD pieces tests/test_oracle.py pins to cv2.SVDecomp / cv2.decomposeEssentialMat).

Tolerances: null direction 1e-9 up to sign (the device accumulates A^T A, the oracle runs a one-sided Jacobi
SVD on A); Euler angles of the winning rotation 2e-6 rad (float storage, as in the reference); translation 2e-6
up to sign (an SVD convention, also with OpenCV)."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu


def _angle_diff(a, b):
    """Largest difference between two Euler triples modulo 2 pi (an angle at +-pi may come out with either sign)."""
    d = (np.asarray(a, np.float64) - np.asarray(b, np.float64) + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d).max()


def _subsets(n, n_samples, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n)[: n // 4] for _ in range(n_samples)]).astype(np.int32)


@pytest.mark.parametrize("n", [64, 2000, 40000])
def test_null_direction_and_normal_matrix(ctx, n):
    b1, b2, r, t, _ = synth.make_two_view(n, seed=n, outlier_frac=0.05)
    idx = _subsets(n, 7, seed=1)
    e, ata = ctx.eight_point_null(b1, b2, idx)
    for s in range(len(idx)):
        A = np.einsum("na,nb->nab", b1[idx[s]], b2[idx[s]]).reshape(-1, 9)
        full = (A.T @ A)[np.triu_indices(9)]
        assert np.allclose(ata[s], full, rtol=1e-12, atol=1e-12)
        e_ref, sv = oracle.eight_point_null(b1, b2, idx[s])
        sgn = np.sign(e[s] @ e_ref)
        assert np.abs(e[s] * sgn - e_ref).max() < 1e-9 and abs(np.linalg.norm(e[s]) - 1) < 1e-12


def test_candidates_match_oracle(ctx):
    b1, b2, r, t, _ = synth.make_two_view(3000, seed=5, outlier_frac=0.05)
    idx = _subsets(3000, 20, seed=2)
    e, _ = ctx.eight_point_null(b1, b2, idx)
    for s in range(len(idx)):
        R1, R2, T, v1, v2 = ctx.essential_to_candidates(e[s])
        o1, o2, oT, w1, w2 = oracle.essential_to_candidates(oracle.eight_point_null(b1, b2, idx[s])[0])
        same = max(_angle_diff(R1, o1), _angle_diff(R2, o2))
        swap = max(_angle_diff(R1, o2), _angle_diff(R2, o1))
        assert min(same, swap) < 2e-6                      # {R1, R2} as a set
        assert sorted([v1, v2]) == sorted([w1, w2])
        assert min(np.abs(T - oT).max(), np.abs(T + oT).max()) < 2e-6


@pytest.mark.parametrize("n,outliers", [(400, 0.0), (8192, 0.05), (8192, 0.3)])
def test_initial_guess_matches_oracle(ctx, n, outliers):
    b1, b2, r, t, _ = synth.make_two_view(n, seed=n + 1, outlier_frac=outliers, rotvec=(0.1, -0.2, 0.3))
    idx = _subsets(n, 80, seed=3)                          # 80 quarter-size subsets, like the reference
    R, T, nc = ctx.initial_guess(b1, b2, idx)
    oR, oT, best, cand = oracle.initial_guess(b1, b2, idx)
    assert nc == len(cand) > 0
    assert np.abs(R - oR).max() < 2e-6
    assert min(np.abs(T - oT).max(), np.abs(T + oT).max()) < 2e-6
    if outliers == 0.0:
        # noise-free: the recovered Euler angles are those of the inverse rotation (the reference negates them, :330)
        from scipy.spatial.transform import Rotation as Rot
        eul = Rot.from_rotvec(r).inv().as_euler("xyz")
        assert np.abs(R - eul).max() < 5e-3


def test_initial_guess_without_valid_candidate_is_an_error(ctx):
    """Both rotation candidates of every subset above 1.57 rad: the reference would index an empty vector (:179);
    the library reports it."""
    from spherical_bundle_adjuster_b200 import SbaError
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(1)
    X1 = synth.unit_rows(rng.standard_normal((400, 3))) * rng.uniform(2, 8, (400, 1))
    R = Rot.from_euler("xyz", [3.0, 0.0, 0.0]).as_matrix()      # a 172 degree roll: |angle| > 1.57 for both candidates
    X2 = X1 @ R.T - np.array([0.3, 0.1, -0.2])
    b1, b2 = synth.unit_rows(X1), synth.unit_rows(X2)
    idx = _subsets(400, 10, seed=4)
    oR, oT, best, cand = oracle.initial_guess(b1, b2, idx)
    if len(cand) == 0:
        with pytest.raises(SbaError):
            ctx.initial_guess(b1, b2, idx)
    else:   # geometry happened to leave a valid twisted-pair member: results must still agree
        R_, T_, nc = ctx.initial_guess(b1, b2, idx)
        assert nc == len(cand) and np.abs(R_ - oR).max() < 2e-6
