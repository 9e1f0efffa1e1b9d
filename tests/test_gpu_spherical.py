"""GPU parity: spherical_surf geometry (spherical_surf.cpp:17-123) through the C ABI, bit-exact against the
oracle -- which tests/test_oracle.py pins to the reference's own spherical_surf.cpp (oracle/_ref)."""
import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu

GEOMS = [(1024, 512), (1920, 960), (2000, 1000), (3840, 1920), (1000, 500), (258, 130)]
PITCHES = [45.0, -45.0, -90.0, 30.5, 0.0]


def test_eular2rot(ctx):
    for th in ([0, 0.3, 0], [0.1, -0.7, 1.2], [0, np.float32(np.pi * 45 / 180), 0], [0, 0, 0]):
        assert np.array_equal(ctx.eular2rot(th), oracle.eular2rot(th))


@pytest.mark.parametrize("w,h", GEOMS)
def test_crop_tables_bit_exact(ctx, w, h):
    for pitch in PITCHES:
        lut, n_patched = ctx.crop_rotated_lut(w, h, pitch)
        ref = oracle.crop_rotated_lut(pitch, w, h)
        assert np.array_equal(lut, ref), (pitch, int((lut != ref).sum()))
        if pitch == 0.0:
            assert n_patched == lut.size          # identity: every coordinate is an integer, all settled on the host


@pytest.mark.parametrize("w,h", [(1024, 512), (3840, 1920), (1000, 500), (258, 130)])
def test_crop_rotated_image(ctx, w, h):
    im = synth.make_erp_image(w, h, seed=w)
    for pitch in (45.0, -45.0, -90.0):
        assert np.array_equal(ctx.crop_rotated_image(im, pitch), oracle.crop_rotated_image(im, pitch))


def test_crop_rotated_image_marks_unmapped_pixels(ctx):
    # pitch -90 on 1024x512 sends one band pixel onto the pole (acos of |z| > 1 is NaN): the reference skips it
    lut, _ = ctx.crop_rotated_lut(1024, 512, -90.0)
    assert (lut < 0).sum() == (oracle.crop_rotated_lut(-90.0, 1024, 512) < 0).sum() >= 1
    im = np.full((512, 1024, 3), 255, np.uint8)
    out = ctx.crop_rotated_image(im, -90.0)
    assert np.array_equal(out == 0, np.repeat((lut < 0)[:, :, None], 3, axis=2))


def test_crop_batched_and_device_mode(ctx):
    import torch
    w, h = 1920, 960
    ims = np.stack([synth.make_erp_image(w, h, seed=s) for s in range(3)])
    want = np.stack([oracle.crop_rotated_image(im, -45.0) for im in ims])
    assert np.array_equal(ctx.crop_rotated_image(ims, -45.0), want)
    dev = ctx.crop_rotated_image(torch.from_numpy(ims).cuda(), -45.0)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)


@pytest.mark.parametrize("w,h", [(1024, 512), (3840, 1920), (1000, 500), (258, 130), (1024, 520)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_four_bands_in_one_gather(ctx, w, h, mode):
    im = synth.make_erp_image(w, h, seed=7)
    ctx.set_remap_kernel(mode)          # per-plan choice, direct, tiled, source-ordered gather (where the geometry allows it)
    try:
        _check_bands(ctx, im, w, h)
    finally:
        ctx.set_remap_kernel(0)


def _check_bands(ctx, im, w, h):
    bands = ctx.spherical_crops(im)
    assert bands.shape == (4, h // 4, w, 3)
    assert np.array_equal(bands[0], oracle.crop_rotated_image(im, 45.0))
    assert np.array_equal(bands[1], im[h * 3 // 8: h * 3 // 8 + h // 4])      # im(roi), spherical_surf.cpp:132,139
    assert np.array_equal(bands[2], oracle.crop_rotated_image(im, -45.0))
    assert np.array_equal(bands[3], oracle.crop_rotated_image(im, -90.0))
    two = ctx.spherical_crops(np.stack([im, im[::-1].copy()]))
    assert np.array_equal(two[0], bands) and np.array_equal(two[1, 1], im[::-1][h * 3 // 8: h * 3 // 8 + h // 4])


@pytest.mark.parametrize("with_table", [False, True])
def test_rotate_keypoints(ctx, with_table):
    w, h = 1920, 960
    rng = np.random.default_rng(3)
    xy = (rng.uniform(0, 1, (5000, 2)) * [w - 1, h / 4 - 1]).astype(np.float32)
    xy[:50] = np.floor(xy[:50])                       # integer keypoints as well
    for pitch in (45.0, -45.0, -90.0):
        if with_table:
            ctx.crop_rotated_lut(w, h, pitch)         # keypoints inside the band then read the crop's own table
        assert np.array_equal(ctx.rotate_keypoints(xy, pitch, w, h), oracle.rotate_keypoints(xy, pitch, w, h))


def test_rotate_pixels_anywhere(ctx):
    w, h = 1024, 512
    rng = np.random.default_rng(4)
    rc = np.stack([rng.integers(0, h, 4000), rng.integers(0, w, 4000)], axis=1).astype(np.int32)
    rc[:4] = [[0, 0], [h // 2, 0], [h // 2, w // 2], [h - 1, w - 1]]
    for pitch in (45.0, -90.0, 12.25):
        assert np.array_equal(ctx.rotate_pixels(rc, pitch, w, h), oracle.rotate_pixels(rc, pitch, w, h))


def test_empty_and_invalid_inputs(ctx):
    from spherical_bundle_adjuster_b200 import SbaError
    assert ctx.rotate_keypoints(np.zeros((0, 2), np.float32), 45.0, 1024, 512).shape == (0, 2)
    assert ctx.rotate_pixels(np.zeros((0, 2), np.int32), 45.0, 1024, 512).shape == (0, 2)
    with pytest.raises(SbaError):
        ctx.crop_rotated_image(np.zeros((3, 8, 3), np.uint8), 45.0)       # fewer than 4 rows: no band
    # keypoints outside the band (and outside the image) go through the direct computation, like the reference
    xy = np.array([[10.0, -40.0], [2000.0, 300.0], [5.5, 127.9]], np.float32)
    assert np.array_equal(ctx.rotate_keypoints(xy, -45.0, 1024, 512), oracle.rotate_keypoints(xy, -45.0, 1024, 512))
