"""Seeded synthetic inputs of the BASELINE shapes.

SURF is non-free and absent from the image, so keypoints/descriptors are synthesised: unit-norm
64-d descriptors with planted correspondences (pure random descriptors give zero survivors at
ratio 0.3), keypoints as bearings pushed through a known rotation and projected onto the cube strip.
NumPy only; no dependency on the CUDA library or on oracle/.
"""
from __future__ import annotations

import numpy as np


def rotvec_to_matrix(r) -> np.ndarray:
    r = np.asarray(r, np.float64)
    th = np.linalg.norm(r)
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def unit_rows(x: np.ndarray) -> np.ndarray:
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def make_descriptors(nq: int, nt: int, dim: int = 64, seed: int = 1, match_frac: float = 0.5, sigma: float = 0.05):
    """Query set A [nq, dim] and train set B [nt, dim] (fp32, unit rows).  A `match_frac` share of the
    queries has a planted partner in B (A row + noise of norm ~sigma, renormalised) at a random train position;
    the rest of B are distractors.  Returns (A, B, truth) with truth[i] = partner index or -1."""
    rng = np.random.default_rng(seed)
    A = unit_rows(rng.standard_normal((nq, dim))).astype(np.float32)
    B = unit_rows(rng.standard_normal((nt, dim))).astype(np.float32)
    n_match = int(min(nq, nt) * match_frac)
    qsel = rng.permutation(nq)[:n_match]
    tsel = rng.permutation(nt)[:n_match]
    noisy = A[qsel].astype(np.float64) + sigma * rng.standard_normal((n_match, dim)) / np.sqrt(dim)
    B[tsel] = unit_rows(noisy).astype(np.float32)
    truth = np.full(nq, -1, np.int64)
    truth[qsel] = tsel
    return A, B, truth


def make_bearings(n: int, rotvec=(0.1, -0.35, 0.6), noise: float = 1e-3, outlier_frac: float = 0.0, seed: int = 2,
                  n_cam: int = 1, rotvec_jitter: float = 0.2):
    """b1 uniform on the sphere, b2 = R_true(cam) b1 + noise (renormalised); a share of outliers gets
    an unrelated b2.  Returns (b1 [n,3] f64, b2 [n,3] f64, cam [n] int32, r_true [n_cam,3])."""
    rng = np.random.default_rng(seed)
    b1 = unit_rows(rng.standard_normal((n, 3)))
    r_true = np.asarray(rotvec, np.float64)[None, :] + (rotvec_jitter * rng.standard_normal((n_cam, 3)) if n_cam > 1 else 0.0)
    cam = rng.integers(0, n_cam, n).astype(np.int32) if n_cam > 1 else np.zeros(n, np.int32)
    b2 = np.empty_like(b1)
    for c in range(n_cam):
        m = cam == c
        if m.any():
            b2[m] = b1[m] @ rotvec_to_matrix(r_true[c]).T
    b2 = unit_rows(b2 + noise * rng.standard_normal((n, 3)))
    n_out = int(n * outlier_frac)
    if n_out:
        sel = rng.permutation(n)[:n_out]
        b2[sel] = unit_rows(rng.standard_normal((n_out, 3)))
    return b1, b2, cam, r_true


def make_two_view(n: int, rotvec=(0.05, -0.1, 0.2), tran=(0.3, 0.1, -0.2), noise: float = 1e-3, outlier_frac: float = 0.05,
                  seed: int = 5, depth_range=(2.0, 8.0)):
    """Points at depth seen from two spherical cameras related by X2 = R X1 - t (the model of the
    reference's functors, spherical_bundle_adjuster.cpp:844-868).  Returns (b1, b2 [n,3] f64 unit
    bearings, r [3], t [3], depths [n,2] = |X1|, |X2|); the first share of matches are outliers."""
    rng = np.random.default_rng(seed)
    X1 = unit_rows(rng.standard_normal((n, 3))) * rng.uniform(depth_range[0], depth_range[1], n)[:, None]
    r, t = np.asarray(rotvec, np.float64), np.asarray(tran, np.float64)
    X2 = X1 @ rotvec_to_matrix(r).T - t
    depths = np.stack([np.linalg.norm(X1, axis=1), np.linalg.norm(X2, axis=1)], axis=1)
    b1 = unit_rows(unit_rows(X1) + noise * rng.standard_normal((n, 3)))
    b2 = unit_rows(X2)
    n_out = int(n * outlier_frac)
    if n_out:
        b2[:n_out] = unit_rows(rng.standard_normal((n_out, 3)))
    return b1, b2, r, t, depths


def make_erp_image(w: int, h: int, seed: int = 3) -> np.ndarray:
    """Seeded uint8 ERP image [h, w, 3] (smooth gradient + noise so neighbouring pixels differ)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = ((xx * 7 + yy * 13) % 251).astype(np.uint8)
    img = np.stack([base, (base[::-1] + 31).astype(np.uint8), ((xx ^ yy) % 256).astype(np.uint8)], axis=2)
    img ^= rng.integers(0, 32, (h, w, 3), dtype=np.uint8)
    return np.ascontiguousarray(img)


def index_image(w: int, h: int) -> np.ndarray:
    """ERP image whose 3 bytes spell the pixel's own linear index (needs w*h < 2**24): remapping it
    reveals the source index of every output pixel."""
    assert w * h < (1 << 24)
    idx = np.arange(w * h, dtype=np.uint32)
    img = np.stack([idx & 255, (idx >> 8) & 255, (idx >> 16) & 255], axis=1).astype(np.uint8)
    return img.reshape(h, w, 3)


def decode_index_image(strip: np.ndarray) -> np.ndarray:
    s = strip.astype(np.int64)
    return (s[..., 0] | (s[..., 1] << 8) | (s[..., 2] << 16)).astype(np.int32)


def bearings_to_strip_xy(b: np.ndarray, cs: int) -> np.ndarray:
    """Project unit bearings onto the 6-face strip (the inverse of equi2cube_surf::cube2equi_pixel);
    returns float32 (x, y) strip keypoint coordinates."""
    b = np.asarray(b, np.float64)
    ax, ay, az = np.abs(b[:, 0]), np.abs(b[:, 1]), np.abs(b[:, 2])
    face = np.where((ay >= ax) & (ay >= az), np.where(b[:, 1] > 0, 0, 2),
                    np.where((ax >= ay) & (ax >= az), np.where(b[:, 0] < 0, 1, 3), np.where(b[:, 2] > 0, 4, 5)))
    i = np.empty(len(b)); j = np.empty(len(b))
    x, y, z = b[:, 0], b[:, 1], b[:, 2]
    m = face == 0; j[m] = cs * (1 - x[m] / y[m]) / 2;      i[m] = cs * (1 - z[m] / y[m]) / 2
    m = face == 1; j[m] = cs * (1 - y[m] / -x[m]) / 2;     i[m] = cs * (1 - z[m] / -x[m]) / 2
    m = face == 2; j[m] = cs * (1 + x[m] / -y[m]) / 2;     i[m] = cs * (1 - z[m] / -y[m]) / 2
    m = face == 3; j[m] = cs * (1 + y[m] / x[m]) / 2;      i[m] = cs * (1 - z[m] / x[m]) / 2
    m = face == 4; i[m] = cs * (1 - x[m] / z[m]) / 2;      j[m] = cs * (1 - y[m] / z[m]) / 2
    m = face == 5; i[m] = cs * (1 + x[m] / -z[m]) / 2;     j[m] = cs * (1 - y[m] / -z[m]) / 2
    j = np.clip(j, 0, np.nextafter(cs, 0))
    xy = np.stack([face * cs + j, np.clip(i, 0, np.nextafter(cs, 0))], axis=1)
    return xy.astype(np.float32)


def make_pair(nq: int = 16384, nt: int = 16384, cs: int = 960, rotvec=(0.1, -0.35, 0.6), seed: int = 4,
              match_frac: float = 0.5, pix_noise: float = 0.0):
    """Keypoints + descriptors of a synthetic ERP pair (BASELINE config 2 stand-in for SURF output).
    Matched train keypoints are the rotated query bearings (plus optional pixel noise); unmatched ones are
    random.  Returns dict(desc1, desc2, key1_xy, key2_xy, truth, r_true)."""
    rng = np.random.default_rng(seed)
    A, B, truth = make_descriptors(nq, nt, 64, seed, match_frac)
    b1 = unit_rows(rng.standard_normal((nq, 3)))
    b2 = unit_rows(rng.standard_normal((nt, 3)))
    R = rotvec_to_matrix(rotvec)
    m = truth >= 0
    b2[truth[m]] = b1[m] @ R.T
    k1 = bearings_to_strip_xy(b1, cs)
    k2 = bearings_to_strip_xy(b2, cs)
    if pix_noise > 0:
        k2 = np.clip(k2 + (pix_noise * rng.standard_normal(k2.shape)).astype(np.float32), 0, None)
    return dict(desc1=A, desc2=B, key1_xy=k1, key2_xy=k2, truth=truth, r_true=np.asarray(rotvec, np.float64))
