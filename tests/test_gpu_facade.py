"""GPU: the C++ drop-in classes (equi2cube, equi2cube_surf, feature_matcher via match_and_lift,
spherical_bundle_adjuster::adjust_rotation) driven from a C++ program, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "build", "facade_demo")


def test_cpp_facade_end_to_end(tmp_path):
    assert os.path.exists(DEMO), "build/facade_demo missing: run __graft_entry__.build()"
    w, h, cs, n1, n2 = 1024, 512, 200, 2500, 2300
    pair = synth.make_pair(n1, n2, cs=cs, seed=31, rotvec=(0.05, 0.2, -0.3))
    im = synth.make_erp_image(w, h, seed=5)
    d = str(tmp_path)
    im.tofile(d + "/im.bin"); pair["desc1"].tofile(d + "/desc1.bin"); pair["desc2"].tofile(d + "/desc2.bin")
    pair["key1_xy"].tofile(d + "/key1.bin"); pair["key2_xy"].tofile(d + "/key2.bin")
    open(d + "/meta.txt", "w").write(f"{w} {h} {cs} {n1} {n2}\n")
    # inputs of the three-stage solve_problem (spherical_bundle_adjuster.cpp:183-217)
    sb1, sb2, r_true, t_true, _ = synth.make_two_view(1500, seed=9)
    sb1, sb2 = sb1.astype(np.float32).astype(np.float64), sb2.astype(np.float32).astype(np.float64)
    r0, t0 = r_true + [0.02, -0.01, 0.03], t_true + [0.03, 0.02, -0.04]
    sb1.tofile(d + "/sp_b1.bin"); sb2.tofile(d + "/sp_b2.bin"); np.concatenate([r0, t0, [1.0]]).tofile(d + "/sp_init.bin")
    # inputs of the spherical_surf post-SURF part: 4 bands x m keypoints/descriptors per side
    m = 300
    rng = np.random.default_rng(12)
    ss_kl = (rng.uniform(0, 1, (4, m, 2)) * [w - 1, h / 4 - 1]).astype(np.float32)
    ss_kr = (rng.uniform(0, 1, (4, m, 2)) * [w - 1, h / 4 - 1]).astype(np.float32)
    ss_dl, ss_dr = synth.make_descriptors(4 * m, 4 * m, 64, seed=13)[:2]
    ss_kl.tofile(d + "/ss_kl.bin"); ss_kr.tofile(d + "/ss_kr.bin"); ss_dl.tofile(d + "/ss_dl.bin"); ss_dr.tofile(d + "/ss_dr.bin")
    r = subprocess.run([DEMO, d], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Ceres Solver Report") == 3

    strip = np.fromfile(d + "/strip.bin", np.uint8).reshape(cs, 6 * cs, 3)
    assert np.array_equal(strip, oracle.equi2cube_all(im, cs))
    back = np.fromfile(d + "/face3.bin", np.uint8).reshape(cs, cs, 3)
    assert np.array_equal(back, oracle.equi2cube_face(im, cs, 3))
    mm = np.fromfile(d + "/matches.bin", np.int32).reshape(-1, 2)
    qi, ti, _ = oracle.match_two_image(pair["desc1"], pair["desc2"], 0.3)
    assert np.array_equal(mm[:, 0], qi) and np.array_equal(mm[:, 1], ti)
    rot = np.fromfile(d + "/rot.bin", np.float64)
    e1 = oracle.cube2equi_points(pair["key1_xy"][qi], cs, w, h); e2 = oracle.cube2equi_points(pair["key2_xy"][ti], cs, w, h)
    b1 = oracle.pixels_to_bearings(e1, w, h).astype(np.float32).astype(np.float64)
    b2 = oracle.pixels_to_bearings(e2, w, h).astype(np.float32).astype(np.float64)
    r_or, s_or = oracle.ba_rot_solve(b1, b2, None, np.zeros((1, 3)))
    assert np.abs(rot[:3] - r_or[0]).max() < 1e-6 and int(rot[3]) == s_or.iterations
    assert np.linalg.norm(rot[:3] - pair["r_true"]) < 1e-4

    sp = np.fromfile(d + "/sp_out.bin", np.float64)
    d_ref, s_d, _ = oracle.ba_d_solve(sb1, sb2, r0, t0, np.full((1500, 2), 1.0))
    r_ref, s_r = oracle.ba_rot_solve(sb1, sb2, None, r0[None], t0, d_ref[0, 0], d_ref[1, 0], 1.0)
    t_ref, s_t = oracle.ba_tran_solve(sb1, sb2, None, r_ref, t0[None], d_ref[0, 0], d_ref[1, 0], 1.0)
    assert np.abs(sp[:3] - r_ref[0]).max() < 1e-6 and np.abs(sp[3:6] - t_ref[0]).max() < 1e-6
    assert int(sp[6]) == s_d.iterations
    assert np.all(np.abs(sp[9:].reshape(-1, 2) - d_ref) <= 1e-7 * np.maximum(1.0, np.abs(d_ref)))

    # spherical_surf facade (spherical_surf.cpp:79-232 minus SURF)
    pitches = (45.0, 0.0, -45.0, -90.0)
    for b, pitch in enumerate(pitches):
        band = np.fromfile(d + f"/ss_band{b}.bin", np.uint8).reshape(h // 4, w, 3)
        want = im[h * 3 // 8: h * 3 // 8 + h // 4] if b == 1 else oracle.crop_rotated_image(im, pitch)
        assert np.array_equal(band, want)
    assert np.array_equal(np.fromfile(d + "/ss_crop45.bin", np.uint8).reshape(h // 4, w, 3), oracle.crop_rotated_image(im, 45.0))
    so = np.fromfile(d + "/ss_out.bin", np.float32)
    assert np.array_equal(so[:2].astype(np.int32), oracle.rotate_pixels(np.array([[h // 2, w // 3]]), -45.0, w, h)[0])
    lifted = []
    for keys in (ss_kl, ss_kr):
        parts = []
        for b, pitch in enumerate(pitches):
            if b == 1:
                k = keys[b].copy(); k[:, 1] = k[:, 1] + np.float32(h * 3 // 8)
            else:
                k = oracle.rotate_keypoints(keys[b], pitch, w, h)
            parts.append(k)
        lifted.append(np.concatenate(parts))
    qi2, ti2, _ = oracle.match_two_image(ss_dl, ss_dr, 0.3)
    rows = so[2:].reshape(-1, 6)
    assert len(rows) == len(qi2) > 0
    assert np.array_equal(rows[:, 0].astype(np.int32), qi2) and np.array_equal(rows[:, 1].astype(np.int32), ti2)
    assert np.array_equal(rows[:, 2:4], lifted[0][qi2]) and np.array_equal(rows[:, 4:6], lifted[1][ti2])

    # initial_guess through the facade: the subsets come from std::random_shuffle on std::rand (seeded 1 by the demo,
    # which is also the C library's state when the reference starts); the same draws are replayed here with libc
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    n_sp, sample_n = 1500, int(1500 * 0.25)
    idx = np.empty((80, sample_n), np.int32)
    for s_ in range(80):
        perm = list(range(n_sp))
        for i in range(1, n_sp):                       # libstdc++'s std::random_shuffle
            j = libc.rand() % (i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        idx[s_] = perm[:sample_n]
    ig = np.fromfile(d + "/ig_out.bin", np.float32)
    oR, oT, best, cand = oracle.initial_guess(sb1, sb2, idx)
    assert np.abs(ig[:3] - oR).max() < 2e-6
    assert min(np.abs(ig[3:] - oT).max(), np.abs(ig[3:] + oT).max()) < 2e-6
