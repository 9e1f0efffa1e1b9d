"""GPU parity: kNN(k=2) + ratio test through the C ABI against cv2.BFMatcher golden vectors and the
oracle.  Indices AND fp32 distances are bit-exact (the CUDA path reproduces OpenCV's accumulation
order); both the exact SIMT kernel and whatever SBA_MATCH_AUTO selects are checked."""
import os

import numpy as np
import pytest

import oracle
from spherical_bundle_adjuster_b200 import MATCH_AUTO, MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16, SbaError, synth

pytestmark = pytest.mark.gpu
ALGOS = [MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16, MATCH_AUTO]
TENSOR_ALGOS = [MATCH_TENSOR, MATCH_TENSOR_FP16]
ERR_BOUND = {MATCH_TENSOR: 4e-5, MATCH_TENSOR_FP16: 1e-3}     # DELTA_COEF / DELTA_COEF_FP16 in matcher_tc.cu


def _match(ctx, q, t, ratio, algo, **kw):
    """Both tensor-core filters cover SURF-64; 128-d descriptors (extended SURF) take the tensor path too (always the fp16
    filter, whichever of the two is asked for)."""
    return ctx.match_two_image(q, t, ratio, algo=algo, **kw)


def _check(m, q, t, ratio=0.3):
    idx, dist = oracle.knn2_l2(q, t)
    qi, ti, dd = oracle.ratio_filter(idx, dist, ratio)
    assert np.array_equal(m.knn_idx, idx)
    assert np.array_equal(m.knn_dist.view(np.uint32), dist.view(np.uint32))
    assert np.array_equal(m.query_idx, qi) and np.array_equal(m.train_idx, ti)
    assert np.array_equal(m.distance.view(np.uint32), dd.view(np.uint32))
    assert (np.diff(m.query_idx) > 0).all()          # survivors in ascending query order


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", ["matcher_64.npz", "matcher_128.npz", "matcher_ragged.npz"])
def test_matcher_matches_cv2_golden(ctx, golden_dir, name, algo):
    g = np.load(os.path.join(golden_dir, name))
    m = _match(ctx, g["q"], g["t"], 0.3, algo, want_knn=True)
    assert np.array_equal(m.knn_idx, g["knn_idx"])
    assert np.array_equal(m.knn_dist.view(np.uint32), g["knn_dist"].view(np.uint32))
    assert np.array_equal(m.query_idx, g["keep"])
    assert np.array_equal(m.train_idx, g["knn_idx"][g["keep"], 0])


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("nq,nt,dim", [(1000, 1000, 64), (777, 1301, 64), (130, 4097, 64), (2048, 300, 128), (5000, 3000, 64), (3000, 4100, 128), (700, 2000, 64), (1290, 900, 128)])
def test_matcher_matches_oracle(ctx, nq, nt, dim, algo):
    A, B, _ = synth.make_descriptors(nq, nt, dim, seed=nq + nt)
    _check(_match(ctx, A, B, 0.3, algo, want_knn=True), A, B)


@pytest.mark.parametrize("algo", ALGOS)
def test_matcher_ties_and_duplicates(ctx, algo):
    A, B, _ = synth.make_descriptors(600, 900, 64, seed=3)
    B[100:140] = B[7]            # forty identical train rows: the two lowest indices must win
    A[5] = B[7]
    B[500] = B[20]; B[20 + 300] = B[20]
    A[9] = A[5]
    _check(ctx.match_two_image(A, B, 0.3, algo=algo, want_knn=True), A, B)
    _check(ctx.match_two_image(A, B, 0.9, algo=algo, want_knn=True), A, B, 0.9)


@pytest.mark.parametrize("algo", ALGOS)
def test_matcher_degenerate_sizes(ctx, algo):
    rng = np.random.default_rng(1)
    q = synth.unit_rows(rng.standard_normal((40, 64))).astype(np.float32)
    m = ctx.match_two_image(q, np.zeros((0, 64), np.float32), algo=algo, want_knn=True)     # M = 0
    assert len(m) == 0 and (m.knn_idx == -1).all() and np.isinf(m.knn_dist).all()
    m = ctx.match_two_image(q, q[:1].copy(), algo=algo, want_knn=True)                        # M = 1
    assert len(m) == 0 and (m.knn_idx[:, 0] == 0).all() and (m.knn_idx[:, 1] == -1).all()
    _check(ctx.match_two_image(q, q[:2].copy(), algo=algo, want_knn=True), q, q[:2])          # M = 2
    m = ctx.match_two_image(np.zeros((0, 64), np.float32), q, algo=algo, want_knn=True)     # N = 0
    assert len(m) == 0


@pytest.mark.parametrize("algo", ALGOS)
def test_matcher_device_tensors(ctx, algo):
    import torch
    A, B, _ = synth.make_descriptors(1500, 1700, 64, seed=9)
    m = ctx.match_two_image(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), 0.3, algo=algo, want_knn=True)
    torch.cuda.synchronize()
    m2 = type(m)(m.query_idx.cpu().numpy(), m.train_idx.cpu().numpy(), m.distance.cpu().numpy(), m.knn_idx.cpu().numpy(), m.knn_dist.cpu().numpy())
    _check(m2, A, B)


@pytest.mark.parametrize("algo", TENSOR_ALGOS)
def test_tensor_path_error_bound_and_fallback(ctx, algo):
    """The re-rank's safety test assumes |approx - exact| <= delta_coef (|a|^2 + max|b|^2) (4e-5 for the bf16 split, 1e-3 for
    the single fp16 product); the kernel reports the largest value it actually saw.  Exact duplicates must route rows
    through the exact fallback."""
    MATCH_TENSOR = algo       # the body below runs once per filter scheme
    A, B, _ = synth.make_descriptors(4096, 4096, 64, seed=21)
    m = ctx.match_two_image(A, B, 0.3, algo=MATCH_TENSOR, want_knn=True)
    st = ctx.match_stats()
    assert st.algo_used == MATCH_TENSOR and st.n_tiles == 32 * 16
    assert 0.0 <= st.max_rel_err < ERR_BOUND[algo] / 4, st.max_rel_err         # observed error, 4x below the assumed bound
    assert st.n_fallback_rows <= (8 if algo == 2 else 64)
    _check(m, A, B)
    B[1000:1040] = B[7]          # 40 identical rows: more ties than candidate chunks can hold
    A[5] = B[7]
    m = ctx.match_two_image(A, B, 0.3, algo=MATCH_TENSOR, want_knn=True)
    assert ctx.match_stats().n_fallback_rows >= 1
    _check(m, A, B)
    # descriptors that are not unit norm (SIFT-like magnitudes): the bound scales with the norms
    A2, B2 = (A * 512).astype(np.float32), (B * 512).astype(np.float32)
    _check(ctx.match_two_image(A2, B2, 0.3, algo=MATCH_TENSOR, want_knn=True), A2, B2)
    assert ctx.match_stats().max_rel_err < ERR_BOUND[algo] / 4


@pytest.mark.parametrize("algo", TENSOR_ALGOS)
@pytest.mark.parametrize("kind", ["all_positive", "dominant_component", "norm_decades", "tiny_and_huge"])
def test_tensor_path_adversarial_error_bound(ctx, kind, algo):
    """Inputs built to stress the filter's error bound (derivation next to DELTA_COEF in matcher_tc.cu): all-positive
    SIFT-like rows (every product of a dot product has the same sign, |a.b| ~ |a||b|), rows with one dominant component,
    norms spanning four decades.  The kNN tables must still equal the exact SIMT kernel bit for bit and the largest
    filter error the kernel saw must stay under the assumed 4e-5 (|a|^2 + max|b|^2)."""
    rng = np.random.default_rng(123)
    nq, nt = 2048, 4096
    if kind == "all_positive":
        A = synth.unit_rows(np.abs(rng.standard_normal((nq, 64)))).astype(np.float32)
        B = synth.unit_rows(np.abs(rng.standard_normal((nt, 64)))).astype(np.float32)
        B[:512] = synth.unit_rows(A[:512].astype(np.float64) + 0.01 * np.abs(rng.standard_normal((512, 64)))).astype(np.float32)
    elif kind == "dominant_component":
        A = (0.02 * rng.standard_normal((nq, 64))).astype(np.float32); A[np.arange(nq), rng.integers(0, 64, nq)] = 1.0
        B = (0.02 * rng.standard_normal((nt, 64))).astype(np.float32); B[np.arange(nt), rng.integers(0, 64, nt)] = 1.0
    elif kind == "norm_decades":
        A = (synth.unit_rows(rng.standard_normal((nq, 64))) * 10.0 ** rng.uniform(-2, 2, (nq, 1))).astype(np.float32)
        B = (synth.unit_rows(rng.standard_normal((nt, 64))) * 10.0 ** rng.uniform(-2, 2, (nt, 1))).astype(np.float32)
    else:
        A = (synth.unit_rows(rng.standard_normal((nq, 64))) * np.where(rng.random((nq, 1)) < 0.5, 1e-3, 1e3)).astype(np.float32)
        B = (synth.unit_rows(rng.standard_normal((nt, 64))) * np.where(rng.random((nt, 1)) < 0.5, 1e-3, 1e3)).astype(np.float32)
    mt = ctx.match_two_image(A, B, 0.3, algo=algo, want_knn=True)
    st = ctx.match_stats()
    ms = ctx.match_two_image(A, B, 0.3, algo=MATCH_SIMT_EXACT, want_knn=True)
    assert st.algo_used == algo
    assert 0.0 <= st.max_rel_err < ERR_BOUND[algo], (kind, st.max_rel_err)
    assert np.array_equal(mt.knn_idx, ms.knn_idx) and np.array_equal(mt.knn_dist.view(np.uint32), ms.knn_dist.view(np.uint32))
    assert np.array_equal(mt.query_idx, ms.query_idx) and np.array_equal(mt.train_idx, ms.train_idx)
    idx, dist = oracle.knn2_l2(A[:64], B)
    assert np.array_equal(mt.knn_idx[:64], idx) and np.array_equal(mt.knn_dist[:64].view(np.uint32), dist.view(np.uint32))
    print(kind, "algo", algo, "fallback rows", st.n_fallback_rows, "max_rel_err", st.max_rel_err)


def test_matcher_full_size_properties(ctx):
    """BASELINE config-2 size (16k x 16k): the oracle needs ~1 min here, so check size-independent
    properties: every planted pair is recovered exactly, AUTO == SIMT bit-for-bit, and a sampled
    subset of rows equals the oracle."""
    A, B, truth = synth.make_descriptors(16384, 16384, 64, seed=2)
    ms = ctx.match_two_image(A, B, 0.3, algo=MATCH_SIMT_EXACT, want_knn=True)
    ma = ctx.match_two_image(A, B, 0.3, algo=MATCH_AUTO, want_knn=True)
    assert ctx.match_stats().algo_used == MATCH_TENSOR
    planted = np.flatnonzero(truth >= 0)
    assert np.array_equal(ms.query_idx, planted) and np.array_equal(ms.train_idx, truth[planted])
    for a, b in [(ms.knn_idx, ma.knn_idx), (ms.query_idx, ma.query_idx), (ms.train_idx, ma.train_idx)]:
        assert np.array_equal(a, b)
    assert np.array_equal(ms.knn_dist.view(np.uint32), ma.knn_dist.view(np.uint32))
    rows = np.random.default_rng(0).choice(16384, 256, replace=False)
    idx, dist = oracle.knn2_l2(A[rows], B)
    assert np.array_equal(ms.knn_idx[rows], idx) and np.array_equal(ms.knn_dist[rows].view(np.uint32), dist.view(np.uint32))


def test_matcher_128d_tensor_path(ctx):
    """128-d descriptors on the tensor path (feature_matcher.cpp:13 -- SURF's extended mode): 16 384 x 16 384 against the exact
    SIMT kernel (full kNN tables, bit for bit) and a row sample against the oracle; AUTO must pick the tensor path."""
    A, B, truth = synth.make_descriptors(16384, 16384, 128, seed=11)
    B[200:230] = B[199]                                     # exact ties -> fallback rows
    ma = ctx.match_two_image(A, B, 0.3, algo=MATCH_AUTO, want_knn=True)
    st = ctx.match_stats()
    assert st.algo_used == MATCH_TENSOR and 0.0 <= st.max_rel_err < ERR_BOUND[MATCH_TENSOR_FP16]
    ms = ctx.match_two_image(A, B, 0.3, algo=MATCH_SIMT_EXACT, want_knn=True)
    assert np.array_equal(ma.knn_idx, ms.knn_idx) and np.array_equal(ma.knn_dist.view(np.uint32), ms.knn_dist.view(np.uint32))
    assert np.array_equal(ma.query_idx, ms.query_idx) and np.array_equal(ma.train_idx, ms.train_idx) and len(ma) > 4000
    rows = np.random.default_rng(1).choice(16384, 128, replace=False)
    idx, dist = oracle.knn2_l2(A[rows], B)
    assert np.array_equal(ma.knn_idx[rows], idx) and np.array_equal(ma.knn_dist[rows].view(np.uint32), dist.view(np.uint32))
    print("128-d: fallback rows", st.n_fallback_rows, "max_rel_err", st.max_rel_err)


@pytest.mark.timeout(900)
def test_matcher_c2_all_rows_against_oracle(ctx):
    """BASELINE config 2, every one of the 16 384 rows: both tensor paths' full kNN tables equal the oracle's (OpenMP, ~1 min)."""
    A, B, _ = synth.make_descriptors(16384, 16384, 64, seed=3)
    oracle.set_threads(os.cpu_count() or 1)
    idx, dist = oracle.knn2_l2(A, B)
    for algo in TENSOR_ALGOS:
        ma = ctx.match_two_image(A, B, 0.3, algo=algo, want_knn=True)
        assert np.array_equal(ma.knn_idx, idx) and np.array_equal(ma.knn_dist.view(np.uint32), dist.view(np.uint32)), algo
        print("algo", algo, "fallback rows", ctx.match_stats().n_fallback_rows, "max_rel_err", ctx.match_stats().max_rel_err)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("algo", TENSOR_ALGOS)
@pytest.mark.parametrize("nq,nt,seed", [(65536, 65536, 1), (50001, 63999, 2)])
def test_matcher_sweep_top_size_against_exact_kernel(ctx, nq, nt, seed, algo):
    """BASELINE config 5's largest size (and a ragged one): tensor path vs the exact SIMT kernel, full kNN tables bit for bit,
    with forty duplicates of one train row (more exact ties than a candidate list holds -> fallback rows)."""
    import torch
    A, B, _ = synth.make_descriptors(nq, nt, 64, seed=seed)
    B[100:140] = B[99]
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    t = ctx.match_two_image(a, b, 0.3, algo=algo, want_knn=True)
    st = ctx.match_stats()
    s = ctx.match_two_image(a, b, 0.3, algo=MATCH_SIMT_EXACT, want_knn=True)
    torch.cuda.synchronize()
    assert torch.equal(t.knn_idx, s.knn_idx) and torch.equal(t.knn_dist.view(torch.int32), s.knn_dist.view(torch.int32))
    assert len(t) == len(s) and torch.equal(t.query_idx, s.query_idx) and torch.equal(t.train_idx, s.train_idx)
    assert st.n_fallback_rows >= 1 and st.max_rel_err < ERR_BOUND[algo]
    print(nq, nt, 'algo', algo, 'fallback rows', st.n_fallback_rows)


@pytest.mark.parametrize("algo", TENSOR_ALGOS)
@pytest.mark.parametrize("n_ctas", [1, 37, 74])
def test_tensor_path_with_fewer_persistent_ctas(ctx, n_ctas, algo):
    """sba_ctx_set_matcher_ctas: the span partition changes, the result does not."""
    A, B, _ = synth.make_descriptors(3000, 5000, 64, seed=77)
    want = oracle.match_two_image(A, B, 0.3)
    ctx.set_matcher_ctas(n_ctas)
    try:
        m = ctx.match_two_image(A, B, 0.3, algo=algo)
        assert ctx.match_stats().n_ctas == n_ctas
    finally:
        ctx.set_matcher_ctas(0)
    assert np.array_equal(m.query_idx, want[0]) and np.array_equal(m.train_idx, want[1])
    assert np.array_equal(m.distance.view(np.uint32), want[2].view(np.uint32))
