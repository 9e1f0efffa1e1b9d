"""Generate the golden fixtures under tests/golden/.  Run in the BUILD container only:

    python tests/golden/make_golden.py

Sources of truth (never the oracle's own C restatement, except where said):
  * matcher_*.npz       cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2) of the installed OpenCV 4.13
                        (the third-party routine the reference's match_two_image calls per
                        north_star) + the ratio loop of feature_matcher.cpp:47-56 in Python.
  * remap_*.npz, remap_lut_hashes.json, cube2equi.npz
                        the REFERENCE's own equi2cube.cpp / equi2cube_surf.cpp, compiled from
                        /root/reference into oracle/_ref (oracle/Makefile) and called here.
  * ba_small.npz        scipy (Rotation.from_rotvec for the rotated points, central differences for
                        the Jacobian) + the oracle's LM result.  Ceres is not installable, so the BA
                        numbers are "parity unpinned" at the Ceres boundary.
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from spherical_bundle_adjuster_b200 import synth  # noqa: E402


def cv2_knn(q, t):
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, 2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), np.inf, np.float32)
    for i, ms in enumerate(knn):
        for k, m in enumerate(ms):
            idx[i, k] = m.trainIdx
            dist[i, k] = m.distance
    keep = [i for i in range(len(q)) if idx[i, 1] >= 0 and dist[i, 0] < np.float32(0.3) * dist[i, 1]]
    return idx, dist, np.array(keep, np.int32)


def matcher_case(name, nq, nt, dim, seed):
    A, B, truth = synth.make_descriptors(nq, nt, dim, seed)
    rng = np.random.default_rng(seed + 100)
    if nt > 8:
        B[nt // 2] = B[1]           # exact duplicate train rows -> distance ties, lower index must win
        B[nt // 2 + 1] = B[1]
        A[0] = B[1]                  # a query identical to a (duplicated) train row: d0 == d1 == 0
        near = B[3].copy(); near[0] = np.nextafter(near[0], np.float32(2))  # near-tie at 1 ulp
        B[nt - 1] = near
        A[1] = (B[3] + np.float32(1e-3) * rng.standard_normal(dim).astype(np.float32))
    idx, dist, keep = cv2_knn(A, B)
    np.savez_compressed(os.path.join(HERE, name), q=A, t=B, knn_idx=idx, knn_dist=dist, keep=keep)
    print(name, "survivors", len(keep))


def lut_from_ref(w, h, cs):
    strip = oracle.ref_equi2cube_all(synth.index_image(w, h), cs)
    return synth.decode_index_image(strip)


def spherical_surf_case():
    """spherical_surf.npz: crop tables, crop samples and rotated keypoints from the REAL reference
    (spherical_surf.cpp compiled into oracle/_ref).  The table is recovered by pushing an index image through
    the reference's crop_rotated_image; a second pass with an all-255 image tells "source pixel 0" from "unwritten"."""
    w, h, seed = 1024, 512, 17
    pitches = np.array([45.0, -45.0, -90.0, 30.5], np.float32)
    im = synth.make_erp_image(w, h, seed=seed)
    rng = np.random.default_rng(5)
    keys = (rng.uniform(0, 1, (300, 2)) * [w - 1, h / 4 - 1]).astype(np.float32)
    shas, samples, rotated = [], [], []
    for p in pitches:
        idx = synth.decode_index_image(oracle.ref_crop_rotated_image(synth.index_image(w, h), float(p))).astype(np.int64)
        # tell "source pixel 0" from "unwritten": an all-ones image stays 0 only where nothing was written
        written = oracle.ref_crop_rotated_image(np.full((h, w, 3), 255, np.uint8), float(p))[:, :, 0] == 255
        lut = np.where(written, idx, -1).astype(np.int32)
        shas.append(hashlib.sha256(lut.tobytes()).hexdigest())
        samples.append(oracle.ref_crop_rotated_image(im, float(p))[::37, ::41])
        rotated.append(oracle.ref_rotate_keypoints(keys, float(p), w, h))
    np.savez_compressed(os.path.join(HERE, "spherical_surf.npz"), w=w, h=h, seed=seed, pitches=pitches, lut_sha256=np.array(shas),
                        crop_samples=np.stack(samples), keys=keys, keys_rotated=np.stack(rotated))


def main():
    assert oracle.ref_available(), "build oracle/_ref first (needs /root/reference)"
    spherical_surf_case()
    if "--only-spherical" in sys.argv:
        return
    matcher_case("matcher_64.npz", 96, 130, 64, 11)
    matcher_case("matcher_128.npz", 70, 65, 128, 12)
    matcher_case("matcher_ragged.npz", 33, 5, 64, 13)

    # remap: small images through the real reference (even and odd cube sizes, h a power of two
    # so the bottom-centre clamp is exercised)
    for name, (w, h, cs) in {"remap_even.npz": (128, 64, 24), "remap_odd.npz": (100, 50, 15)}.items():
        im = synth.make_erp_image(w, h, seed=21)
        strip = oracle.ref_equi2cube_all(im, cs)
        faces = np.stack([oracle.ref_equi2cube_face(im, cs, f) for f in range(6)])
        np.savez_compressed(os.path.join(HERE, name), im=im, cs=cs, strip=strip, faces=faces)
    hashes = {}
    for (w, h, cs) in [(2048, 1024, 512), (2048, 1024, 600), (3840, 1920, 960), (512, 256, 128), (100, 50, 15)]:
        lut = lut_from_ref(w, h, cs)
        # the reference reads one row past the image at the bottom centre (h a power of two); the
        # padded row replicates row h-1, so decode yields row h: fold it back to the documented clamp
        lut = np.where(lut >= w * h, lut - w, lut).astype(np.int32)
        hashes[f"{w}x{h}_cs{cs}"] = dict(sha256=hashlib.sha256(lut.tobytes()).hexdigest(), sum=int(lut.astype(np.int64).sum()))
    json.dump(hashes, open(os.path.join(HERE, "remap_lut_hashes.json"), "w"), indent=1)

    rng = np.random.default_rng(31)
    cs, w, h = 600, 2048, 1024
    xy = np.stack([rng.uniform(0, 6 * cs, 512), rng.uniform(0, cs, 512)], 1).astype(np.float32)
    xy[:12, 0] = np.arange(12) * (cs / 2.0)   # face boundaries and centres
    xy[12:16] = [[-3.0, 10.0], [6 * cs + 2.0, 5.0], [cs, 0.0], [5 * cs, cs]]
    np.savez_compressed(os.path.join(HERE, "cube2equi.npz"), xy=xy, cs=cs, w=w, h=h, out=oracle.ref_cube2equi_points(xy, cs, w, h))

    # BA: values from scipy, LM trajectory from the oracle
    n = 64
    b1, b2, cam, r_true = synth.make_bearings(n, noise=2e-3, outlier_frac=0.1, seed=41)
    b1 = b1.astype(np.float32).astype(np.float64); b2 = b2.astype(np.float32).astype(np.float64)
    r = np.array([0.3, 0.2, -0.1]); t = np.array([0.01, -0.02, 0.03]); d1, d2 = 1.1, 0.9
    rot = lambda rv: Rotation.from_rotvec(rv).apply(b1 * d1)
    res = b2 * d2 - (rot(r) - t)
    eps = 1e-6
    jac = np.empty((n, 3, 3))
    for k in range(3):
        e = np.zeros(3); e[k] = eps
        jac[:, :, k] = -(rot(r + e) - rot(r - e)) / (2 * eps)
    s = (res ** 2).sum(1)
    rho = np.where(s <= 1, s, 2 * np.sqrt(s) - 1); w8 = np.where(s <= 1, 1.0, 1 / np.sqrt(np.maximum(s, 1e-300)))
    H = np.einsum("n,nak,nal->kl", w8, jac, jac); g = np.einsum("n,nak,na->k", w8, jac, res)
    r_fin, summ = oracle.ba_rot_solve(b1, b2, None, np.zeros((1, 3)))
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), b1=b1, b2=b2, r=r, t=t, d1=d1, d2=d2, res=res, jac=jac,
                        H=np.array([H[0, 0], H[0, 1], H[0, 2], H[1, 1], H[1, 2], H[2, 2]]), g=g, cost=0.5 * rho.sum(),
                        r_solved=r_fin, iterations=summ.iterations, final_cost=summ.final_cost, r_true=r_true)
    print("golden written")


if __name__ == "__main__":
    main()
