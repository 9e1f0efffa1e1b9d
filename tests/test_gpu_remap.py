"""GPU parity: equi2cube gather, cube2equi keypoints and pixel->bearing through the C ABI against
the oracle and the committed golden fixtures.  Byte and index outputs are bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["remap_even.npz", "remap_odd.npz"])
def test_remap_matches_reference_golden(ctx, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    cs = int(g["cs"])
    assert np.array_equal(ctx.equi2cube(g["im"], cs), g["strip"])
    for f in range(6):
        assert np.array_equal(ctx.equi2cube_face(g["im"], cs, f), g["faces"][f])


@pytest.mark.parametrize("geom", ["2048x1024_cs512", "2048x1024_cs600", "3840x1920_cs960", "512x256_cs128", "100x50_cs15"])
def test_lut_bit_exact_against_reference_hash(ctx, golden_dir, geom):
    """The index table of every BASELINE geometry equals the one the reference's own code produces
    (hash of the table decoded from an index image run through oracle/_ref at fixture time)."""
    h = json.load(open(os.path.join(golden_dir, "remap_lut_hashes.json")))[geom]
    wh, cs = geom.split("_cs")
    w, hh = map(int, wh.split("x"))
    lut = ctx.equi2cube_lut(w, hh, int(cs))
    assert int(lut.astype(np.int64).sum()) == h["sum"]
    assert hashlib.sha256(lut.tobytes()).hexdigest() == h["sha256"]


@pytest.mark.parametrize("w,h,cs", [(512, 256, 128), (640, 320, 77), (2048, 1024, 600)])
def test_remap_matches_oracle(ctx, w, h, cs):
    im = synth.make_erp_image(w, h, seed=w + cs)
    assert np.array_equal(ctx.equi2cube(im, cs), oracle.equi2cube_all(im, cs))


def test_remap_batch_and_device_tensors(ctx):
    import torch
    w, h, cs, n = 512, 256, 96, 5
    ims = np.stack([synth.make_erp_image(w, h, seed=s) for s in range(n)])
    ref = np.stack([oracle.equi2cube_all(im, cs) for im in ims])
    assert np.array_equal(ctx.equi2cube(ims, cs), ref)                      # host batch
    out = ctx.equi2cube(torch.from_numpy(ims).cuda(), cs)                   # device batch, zero copy
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    # odd cube size with a batch: per-image bases are not 4-byte aligned -> generic path
    ref_odd = np.stack([oracle.equi2cube_all(im, 33) for im in ims])
    assert np.array_equal(ctx.equi2cube(ims, 33), ref_odd)


def test_remap_full_size_index_property(ctx):
    """BASELINE config-2 size: remapping the index image must reproduce the table itself."""
    w, h, cs = 3840, 1920, 960
    strip = ctx.equi2cube(synth.index_image(w, h), cs)
    assert np.array_equal(synth.decode_index_image(strip), ctx.equi2cube_lut(w, h, cs))


def test_cube2equi_points(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "cube2equi.npz"))
    out = ctx.cube2equi_points(g["xy"], int(g["cs"]), int(g["w"]), int(g["h"]))
    # fp64 math on both sides, fp32 result: CUDA's acos/atan2 may differ from glibc in the last fp64
    # ulp, which can move an fp32 rounding only on an exact tie -> tolerance 1 fp32 ulp
    assert np.allclose(out, g["out"], rtol=2e-7, atol=1e-6)
    assert (out.view(np.uint32) == g["out"].view(np.uint32)).mean() > 0.99
    rng = np.random.default_rng(5)
    xy = np.stack([rng.uniform(0, 6 * 960, 20000), rng.uniform(0, 960, 20000)], 1).astype(np.float32)
    assert np.allclose(ctx.cube2equi_points(xy, 960, 3840, 1920), oracle.cube2equi_points(xy, 960, 3840, 1920), rtol=2e-7, atol=1e-6)


def test_pixels_to_bearings(ctx):
    rng = np.random.default_rng(6)
    xy = np.stack([rng.uniform(0, 3840, 10000), rng.uniform(0, 1920, 10000)], 1).astype(np.float32)
    b32, b64 = ctx.pixels_to_bearings(xy, 3840, 1920, want_f64=True)
    ref = oracle.pixels_to_bearings(xy, 3840, 1920)
    assert np.abs(b64 - ref).max() < 1e-14          # fp64 tolerance: CUDA sincos <= 2 ulp
    assert np.abs(b32[:, :3] - ref).max() < 1e-7 and (b32[:, 3] == 0).all()
    assert ctx.pixels_to_bearings(np.zeros((0, 2), np.float32), 64, 32).shape == (0, 4)   # empty input


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_all_gather_kernels_give_the_same_strips(ctx, mode):
    """Direct, tiled (bulk-copy staged) and source-ordered gather, forced in turn: bit-identical strips on a geometry with
    fallback tiles (poles of the top and bottom faces) and a batch of frames."""
    w, h, cs = 2048, 1024, 512
    ims = np.stack([synth.make_erp_image(w, h, seed=s) for s in range(3)])
    want = np.stack([oracle.equi2cube_all(im, cs) for im in ims])
    info = ctx.remap_plan_info(w, h, cs)
    assert info["tiled_available"] and 0 < info["n_fallback_tiles"] < info["n_tiles"]
    assert info["sorted"]["available"] and info["sorted"]["n_fallback_tiles"] == 0
    ctx.set_remap_kernel(mode)
    try:
        assert np.array_equal(ctx.equi2cube(ims, cs), want)
        assert np.array_equal(ctx.equi2cube(ims[1], cs), want[1])
    finally:
        ctx.set_remap_kernel(0)


def test_source_ordered_gather_odd_geometries(ctx):
    """Source-ordered gather: a strip width that is no multiple of the 64-pixel tile (cube 80: the form is not built and the
    direct gather runs), a small cube, and the tile-local fallback -- 64-pixel faces over an 8K source make a tile's source
    span exceed the 21-bit offsets of an entry."""
    for (w, h, cs) in [(512, 256, 80), (640, 320, 192)]:
        im = synth.make_erp_image(w, h, seed=3)
        ctx.set_remap_kernel(3)
        try:
            assert ctx.remap_plan_info(w, h, cs)["sorted"]["available"] == ((6 * cs) % 64 == 0)
            assert np.array_equal(ctx.equi2cube(im, cs), oracle.equi2cube_all(im, cs))
        finally:
            ctx.set_remap_kernel(0)
    w, h, cs = 8192, 4096, 64          # 64-pixel faces over an 8K source: every tile spans hundreds of source rows
    im = synth.make_erp_image(w, h, seed=5)
    ctx.set_remap_kernel(3)
    try:
        info = ctx.remap_plan_info(w, h, cs)["sorted"]
        assert info["available"] and info["n_fallback_tiles"] > 0
        assert np.array_equal(ctx.equi2cube(im, cs), oracle.equi2cube_all(im, cs))
    finally:
        ctx.set_remap_kernel(0)
