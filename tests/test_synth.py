"""CPU tests of the synthetic-input generator used by the GPU tests and bench.py."""
import numpy as np

import oracle
from spherical_bundle_adjuster_b200 import synth


def test_planted_descriptors_survive_ratio_test():
    A, B, truth = synth.make_descriptors(1500, 1200, 64, seed=2)
    qi, ti, _ = oracle.match_two_image(A, B, 0.3)
    planted = np.flatnonzero(truth >= 0)
    assert np.array_equal(qi, planted) and np.array_equal(ti, truth[planted])


def test_strip_projection_roundtrip():
    rng = np.random.default_rng(0)
    b = synth.unit_rows(rng.standard_normal((5000, 3)))
    cs, w, h = 600, 2048, 1024
    xy = synth.bearings_to_strip_xy(b, cs)
    assert (xy[:, 0] >= 0).all() and (xy[:, 0] < 6 * cs).all() and (xy[:, 1] >= 0).all() and (xy[:, 1] < cs).all()
    back = oracle.pixels_to_bearings(oracle.cube2equi_points(xy, cs, w, h), w, h)
    assert np.linalg.norm(back - b, axis=1).max() < 5e-6


def test_index_image_roundtrip():
    w, h, cs = 128, 64, 16
    strip = oracle.equi2cube_all(synth.index_image(w, h), cs)
    lut, _ = oracle.equi2cube_lut(cs, w, h)
    assert np.array_equal(synth.decode_index_image(strip), lut)


def test_rotation_recovered_within_two_degrees_like_reference_harness():
    # the reference's own acceptance idea (test/feature_test.cpp:36-62,:208): angle between the
    # rotated left bearing and the right bearing <= 2 degrees for inliers
    p = synth.make_pair(800, 800, cs=128, seed=8)
    m = p["truth"] >= 0
    e1 = oracle.cube2equi_points(p["key1_xy"][m], 128, 512, 256)
    e2 = oracle.cube2equi_points(p["key2_xy"][p["truth"][m]], 128, 512, 256)
    b1, b2 = oracle.pixels_to_bearings(e1, 512, 256), oracle.pixels_to_bearings(e2, 512, 256)
    ang = np.degrees(np.arccos(np.clip(((b1 @ synth.rotvec_to_matrix(p["r_true"]).T) * b2).sum(1), -1, 1)))
    assert (ang <= 2.0).all()
