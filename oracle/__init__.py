"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

ctypes bindings over ``oracle/libsba_oracle.so`` (the plain-C restatement in ``sba_oracle.c``) and,
when it was built, ``oracle/_ref/libsba_ref.so`` (the reference's own ``equi2cube.cpp`` /
``equi2cube_surf.cpp`` compiled from ``/root/reference`` against the cv type shim).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product package ``spherical_bundle_adjuster_b200`` never
does; it fails loudly when its CUDA library is missing instead of falling back to anything here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsba_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libsba_ref.so")


def build(verbose: bool = False) -> None:
    """Compile the oracle (and ``_ref`` when /root/reference is present) with oracle/Makefile."""
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    r = subprocess.run(["make", "-C", _HERE, "all"], env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _load():
    if not os.path.exists(_LIB_PATH):
        build()
    lib = C.CDLL(_LIB_PATH)
    i32p, f32p, f64p, u8p = (C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8))
    lib.orc_max_threads.restype = C.c_int
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_equi2cube_src_index.restype = C.c_int32
    lib.orc_equi2cube_src_index.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)]
    lib.orc_equi2cube_lut.restype = C.c_int
    lib.orc_equi2cube_lut.argtypes = [C.c_int, C.c_int, C.c_int, i32p]
    lib.orc_equi2cube_face.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    lib.orc_equi2cube_all.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
    lib.orc_cube2equi_points.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    lib.orc_pixels_to_bearings.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f64p]
    lib.orc_l2sqr_opencv.restype = C.c_float
    lib.orc_l2sqr_opencv.argtypes = [f32p, f32p, C.c_int]
    lib.orc_knn2_l2.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p, f32p]
    lib.orc_ratio_filter.restype = C.c_int
    lib.orc_ratio_filter.argtypes = [i32p, f32p, C.c_int, C.c_float, i32p, i32p, f32p]
    lib.orc_angle_axis_rotate_point.argtypes = [f64p, f64p, f64p]
    lib.orc_ba_rot_functor.argtypes = [f64p, f64p, f64p, f64p, C.c_double, C.c_double, f64p, f64p]
    lib.orc_ba_rot_eval.argtypes = [f64p, f64p, i32p, C.c_int, f64p, C.c_int, f64p, C.c_double, C.c_double, C.c_double,
                                    f64p, f64p, f64p, f64p, f64p]
    lib.orc_ba_rot_solve.argtypes = [f64p, f64p, i32p, C.c_int, f64p, C.c_int, f64p, C.c_double, C.c_double, C.c_double,
                                     C.c_int, C.c_void_p]
    lib.orc_ba_tran_eval.argtypes = [f64p, f64p, i32p, C.c_int, f64p, f64p, C.c_int, C.c_double, C.c_double, C.c_double, f64p, f64p, f64p, f64p]
    lib.orc_ba_tran_solve.argtypes = [f64p, f64p, i32p, C.c_int, f64p, f64p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
    lib.orc_ba_d_functor.argtypes = [f64p, f64p, f64p, f64p, f64p, C.c_double, C.c_double, f64p, f64p]
    lib.orc_ba_d_solve.argtypes = [f64p, f64p, C.c_int, f64p, f64p, f64p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.orc_eular2rot.argtypes = [f32p, f64p]
    lib.orc_rotate_pixel.argtypes = [C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.orc_crop_rotated_lut.argtypes = [C.c_float, C.c_int, C.c_int, i32p]
    lib.orc_crop_rotated_image.argtypes = [u8p, C.c_int, C.c_int, C.c_float, u8p]
    lib.orc_rotate_keypoints.argtypes = [C.c_float, f32p, C.c_int, C.c_int, C.c_int]
    lib.orc_essential_to_candidates.argtypes = [f64p, f32p, f32p, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.orc_eight_point_null.argtypes = [f64p, f64p, i32p, C.c_int, f64p, f64p]
    lib.orc_vote_rotation.restype = C.c_int
    lib.orc_vote_rotation.argtypes = [f32p, C.c_int]
    lib.orc_initial_guess.restype = C.c_int
    lib.orc_initial_guess.argtypes = [f64p, f64p, i32p, C.c_int, C.c_int, f32p, f32p, f32p, C.POINTER(C.c_int)]
    lib.orc_ls_next_step.restype = C.c_double
    lib.orc_ls_next_step.argtypes = [C.c_double] * 5 + [C.c_int] + [C.c_double] * 5
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


# ------------------------------------------------------------------ remap
def equi2cube_lut(cs: int, w: int, h: int):
    lut = np.empty((cs, 6 * cs), np.int32)
    nclamp = lib().orc_equi2cube_lut(cs, w, h, _p(lut, C.c_int32))
    return lut, nclamp


def equi2cube_all(im: np.ndarray, cs: int) -> np.ndarray:
    im = np.ascontiguousarray(im, np.uint8)
    h, w, _ = im.shape
    out = np.empty((cs, 6 * cs, 3), np.uint8)
    lib().orc_equi2cube_all(_p(im, C.c_uint8), w, h, cs, _p(out, C.c_uint8))
    return out


def equi2cube_face(im: np.ndarray, cs: int, face: int) -> np.ndarray:
    im = np.ascontiguousarray(im, np.uint8)
    h, w, _ = im.shape
    out = np.empty((cs, cs, 3), np.uint8)
    lib().orc_equi2cube_face(_p(im, C.c_uint8), w, h, cs, face, _p(out, C.c_uint8))
    return out


def cube2equi_points(xy: np.ndarray, cs: int, w: int, h: int) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(xy)
    lib().orc_cube2equi_points(_p(xy, C.c_float), len(xy), cs, w, h, _p(out, C.c_float))
    return out


def pixels_to_bearings(xy: np.ndarray, w: int, h: int) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty((len(xy), 3), np.float64)
    lib().orc_pixels_to_bearings(_p(xy, C.c_float), len(xy), w, h, _p(out, C.c_double))
    return out


# ------------------------------------------------------------------ matcher
def knn2_l2(q: np.ndarray, t: np.ndarray):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    nq, dim = q.shape if q.ndim == 2 else (0, t.shape[1])
    nt = t.shape[0]
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.float32)
    lib().orc_knn2_l2(_p(q, C.c_float), nq, _p(t, C.c_float), nt, dim, _p(idx, C.c_int32), _p(dist, C.c_float))
    return idx, dist


def ratio_filter(idx: np.ndarray, dist: np.ndarray, ratio: float = 0.3):
    nq = idx.shape[0]
    qi = np.empty(nq, np.int32)
    ti = np.empty(nq, np.int32)
    d = np.empty(nq, np.float32)
    n = lib().orc_ratio_filter(_p(np.ascontiguousarray(idx), C.c_int32), _p(np.ascontiguousarray(dist), C.c_float), nq,
                               C.c_float(ratio), _p(qi, C.c_int32), _p(ti, C.c_int32), _p(d, C.c_float))
    return qi[:n].copy(), ti[:n].copy(), d[:n].copy()


def match_two_image(q: np.ndarray, t: np.ndarray, ratio: float = 0.3):
    """feature_matcher::match_two_image (feature_matcher.cpp:42-59) with BFMatcher semantics."""
    idx, dist = knn2_l2(q, t)
    return ratio_filter(idx, dist, ratio)


# ------------------------------------------------------------------ bundle adjustment
@dataclass
class LMSummary:
    iterations: int
    num_successful: int
    termination: int
    initial_cost: float
    final_cost: float
    final_radius: float


class _CSummary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_successful", C.c_int), ("termination", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double)]


def _ba_args(b1, b2, cam, r, t):
    b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
    b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(r, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float64).reshape(3)
    cam = None if cam is None else np.ascontiguousarray(cam, np.int32)
    return b1, b2, cam, r, t


def ba_rot_eval(b1, b2, cam, r, t=(0, 0, 0), d1=1.0, d2=1.0, huber=1.0, want_res=True, want_jac=True):
    b1, b2, cam, r, t = _ba_args(b1, b2, cam, r, t)
    n, n_cam = len(b1), len(r)
    res = np.empty((n, 3)) if want_res else None
    jac = np.empty((n, 3, 3)) if want_jac else None
    H = np.empty((n_cam, 6)); g = np.empty((n_cam, 3)); cost = np.empty(n_cam)
    lib().orc_ba_rot_eval(_p(b1, C.c_double), _p(b2, C.c_double), _p(cam, C.c_int32), n, _p(r, C.c_double), n_cam,
                          _p(t, C.c_double), d1, d2, huber, _p(res, C.c_double), _p(jac, C.c_double),
                          _p(H, C.c_double), _p(g, C.c_double), _p(cost, C.c_double))
    return res, jac, H, g, cost


def ba_rot_solve(b1, b2, cam, r0, t=(0, 0, 0), d1=1.0, d2=1.0, huber=1.0, max_iter=50):
    b1, b2, cam, r, t = _ba_args(b1, b2, cam, r0, t)
    r = r.copy()
    s = _CSummary()
    lib().orc_ba_rot_solve(_p(b1, C.c_double), _p(b2, C.c_double), _p(cam, C.c_int32), len(b1), _p(r, C.c_double), len(r),
                           _p(t, C.c_double), d1, d2, huber, max_iter, C.byref(s))
    return r, LMSummary(s.iterations, s.num_successful, s.termination, s.initial_cost, s.final_cost, s.final_radius)


def ba_tran_eval(b1, b2, cam, r, tv, d1=1.0, d2=1.0, huber=1.0):
    """Translation-only functor (spherical_bundle_adjuster.cpp:948-1002): r fixed, tv [n_cam x 3] free."""
    b1, b2, cam, r, _ = _ba_args(b1, b2, cam, r, (0, 0, 0))
    tv = np.ascontiguousarray(tv, np.float64).reshape(-1, 3)
    n, n_cam = len(b1), len(r)
    res = np.empty((n, 3)); H = np.empty((n_cam, 6)); g = np.empty((n_cam, 3)); cost = np.empty(n_cam)
    lib().orc_ba_tran_eval(_p(b1, C.c_double), _p(b2, C.c_double), _p(cam, C.c_int32), n, _p(r, C.c_double), _p(tv, C.c_double), n_cam,
                           d1, d2, huber, _p(res, C.c_double), _p(H, C.c_double), _p(g, C.c_double), _p(cost, C.c_double))
    return res, H, g, cost


def ba_tran_solve(b1, b2, cam, r, t0, d1=1.0, d2=1.0, huber=1.0, max_iter=50):
    b1, b2, cam, r, _ = _ba_args(b1, b2, cam, r, (0, 0, 0))
    tv = np.ascontiguousarray(t0, np.float64).reshape(-1, 3).copy()
    s = _CSummary()
    lib().orc_ba_tran_solve(_p(b1, C.c_double), _p(b2, C.c_double), _p(cam, C.c_int32), len(b1), _p(r, C.c_double), _p(tv, C.c_double), len(r),
                            d1, d2, huber, max_iter, C.byref(s))
    return tv, LMSummary(s.iterations, s.num_successful, s.termination, s.initial_cost, s.final_cost, s.final_radius)


def ba_d_functor(b1, b2, r, t, d, lam=1.0, c=1.0):
    """One depth-only residual block: (res[5], J[5, 2]).  spherical_bundle_adjuster.cpp:1005-1032."""
    a = [np.ascontiguousarray(v, np.float64).reshape(-1) for v in (b1, b2, r, t, d)]
    res, J = np.zeros(5), np.zeros((5, 2))
    lib().orc_ba_d_functor(*[_p(v, C.c_double) for v in a], lam, c, _p(res, C.c_double), _p(J, C.c_double))
    return res, J


def ba_d_solve(b1, b2, r, t, d0, lam=1.0, c=1.0, max_iter=50):
    """Depth-only stage of solve_problem (:196-197): returns (d [n, 2], LMSummary, line-search trials beyond alpha=1)."""
    b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
    b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(r, np.float64).reshape(3)
    t = np.ascontiguousarray(t, np.float64).reshape(3)
    d = np.ascontiguousarray(d0, np.float64).reshape(-1, 2).copy()
    assert len(d) == len(b1) == len(b2)
    s, nls = _CSummary(), C.c_int(0)
    lib().orc_ba_d_solve(_p(b1, C.c_double), _p(b2, C.c_double), len(b1), _p(r, C.c_double), _p(t, C.c_double), _p(d, C.c_double),
                         lam, c, max_iter, C.byref(s), C.byref(nls))
    return d, LMSummary(s.iterations, s.num_successful, s.termination, s.initial_cost, s.final_cost, s.final_radius), nls.value


def ls_next_step(f0, g0, cur, prev=None, min_step=None, max_step=None):
    """Ceres' Armijo step-size interpolation (CUBIC): cur/prev = (x, value, gradient)."""
    xp, fp, gp = prev if prev is not None else (0.0, 0.0, 0.0)
    xc, fc, gc = cur
    lo = 1e-3 * xc if min_step is None else min_step
    hi = 0.6 * xc if max_step is None else max_step
    return lib().orc_ls_next_step(f0, g0, xp, fp, gp, int(prev is not None), xc, fc, gc, lo, hi)


# ------------------------------------------------------------------ initial guess (8-point voting)
def eight_point_null(b1, b2, idx=None):
    """Null direction e [9] (sign arbitrary) and singular values [9] of the epipolar system of one subset."""
    b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
    b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
    idx = None if idx is None else np.ascontiguousarray(idx, np.int32)
    n = len(b1) if idx is None else len(idx)
    e, sv = np.zeros(9), np.zeros(9)
    lib().orc_eight_point_null(_p(b1, C.c_double), _p(b2, C.c_double), _p(idx, C.c_int32), n, _p(e, C.c_double), _p(sv, C.c_double))
    return e, sv


def essential_to_candidates(e):
    """(R1_vec, R2_vec, T_vec float32 [3], R1_valid, R2_valid) of eight_point_estimation from the null direction."""
    e = np.ascontiguousarray(e, np.float64).reshape(9)
    R1, R2, T = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    v1, v2 = C.c_int(0), C.c_int(0)
    lib().orc_essential_to_candidates(_p(e, C.c_double), _p(R1, C.c_float), _p(R2, C.c_float), _p(T, C.c_float), C.byref(v1), C.byref(v2))
    return R1, R2, T, bool(v1.value), bool(v2.value)


def initial_guess(b1, b2, idx):
    """initial_guess with explicit subsets idx [n_samples, sample_n]: (R_vec [3] f32, T_vec [3] f32, winner, candidates [r, 3])."""
    b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
    b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, np.int32)
    ns, sn = idx.shape
    R, T, cand, nc = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros((2 * ns, 3), np.float32), C.c_int(0)
    best = lib().orc_initial_guess(_p(b1, C.c_double), _p(b2, C.c_double), _p(idx, C.c_int32), ns, sn, _p(R, C.c_float), _p(T, C.c_float),
                                   _p(cand, C.c_float), C.byref(nc))
    return R, T, best, cand[:nc.value]


# ------------------------------------------------------------------ spherical_surf geometry
def eular2rot(theta) -> np.ndarray:
    th = np.ascontiguousarray(theta, np.float32).reshape(3)
    R = np.zeros(9)
    lib().orc_eular2rot(_p(th, C.c_float), _p(R, C.c_double))
    return R.reshape(3, 3)


def pitch_rotation(pitch_deg: float) -> np.ndarray:
    """eular2rot(Vec3f(0, RAD(pitch), 0)) as crop_rotated_image / rotate_keypoint build it."""
    return eular2rot([0.0, np.float32(np.pi * float(np.float32(pitch_deg)) / 180.0), 0.0])


def rotate_pixels(rc: np.ndarray, pitch_deg: float, w: int, h: int) -> np.ndarray:
    rc = np.ascontiguousarray(rc, np.int32).reshape(-1, 2)
    R = np.ascontiguousarray(pitch_rotation(pitch_deg).reshape(-1))
    out = np.empty_like(rc)
    a, b = C.c_int(0), C.c_int(0)
    for k in range(len(rc)):
        lib().orc_rotate_pixel(int(rc[k, 0]), int(rc[k, 1]), _p(R, C.c_double), w, h, C.byref(a), C.byref(b))
        out[k] = (a.value, b.value)
    return out


def crop_rotated_lut(pitch_deg: float, w: int, h: int) -> np.ndarray:
    lut = np.empty((h // 4, w), np.int32)
    lib().orc_crop_rotated_lut(pitch_deg, w, h, _p(lut, C.c_int32))
    return lut


def crop_rotated_image(im: np.ndarray, pitch_deg: float) -> np.ndarray:
    im = np.ascontiguousarray(im, np.uint8)
    h, w, _ = im.shape
    out = np.empty((h // 4, w, 3), np.uint8)
    lib().orc_crop_rotated_image(_p(im, C.c_uint8), w, h, pitch_deg, _p(out, C.c_uint8))
    return out


def rotate_keypoints(xy: np.ndarray, pitch_inv_deg: float, w: int, h: int) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2).copy()
    lib().orc_rotate_keypoints(pitch_inv_deg, _p(xy, C.c_float), len(xy), w, h)
    return xy


# ------------------------------------------------------------------ the real reference (oracle/_ref)
def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


_ref = None


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libsba_ref.so not built (needs /root/reference at build time)")
        r = C.CDLL(_REF_PATH)
        u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        r.ref_eular2rot.argtypes = [f32p, C.POINTER(C.c_double)]
        r.ref_rotate_pixels.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int)]
        r.ref_crop_rotated_image.argtypes = [u8p, C.c_int, C.c_int, C.c_float, C.c_int, u8p]
        r.ref_rotate_keypoints.argtypes = [C.c_float, f32p, C.c_int, C.c_int, C.c_int]
        r.ref_equi2cube_all.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        r.ref_equi2cube_face.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        r.ref_cube2equi_points.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _ref = r
    return _ref


def _pad_row(im: np.ndarray) -> np.ndarray:
    """One replicated padding row so the reference's unchecked bottom-centre read stays in bounds."""
    return np.ascontiguousarray(np.concatenate([im, im[-1:]], axis=0), np.uint8)


def ref_equi2cube_all(im: np.ndarray, cs: int, nthreads: int = 0) -> np.ndarray:
    h, w, _ = im.shape
    imp = _pad_row(im)
    out = np.empty((cs, 6 * cs, 3), np.uint8)
    ref().ref_equi2cube_all(_p(imp, C.c_uint8), w, h, cs, nthreads or (os.cpu_count() or 1), _p(out, C.c_uint8))
    return out


def ref_equi2cube_face(im: np.ndarray, cs: int, face: int) -> np.ndarray:
    h, w, _ = im.shape
    imp = _pad_row(im)
    out = np.empty((cs, cs, 3), np.uint8)
    ref().ref_equi2cube_face(_p(imp, C.c_uint8), w, h, cs, face, _p(out, C.c_uint8))
    return out


def ref_cube2equi_points(xy: np.ndarray, cs: int, w: int, h: int) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(xy)
    ref().ref_cube2equi_points(_p(xy, C.c_float), len(xy), cs, w, h, _p(out, C.c_float))
    return out


def ref_eular2rot(theta) -> np.ndarray:
    th = np.ascontiguousarray(theta, np.float32).reshape(3)
    R = np.zeros(9)
    ref().ref_eular2rot(_p(th, C.c_float), _p(R, C.c_double))
    return R.reshape(3, 3)


def ref_rotate_pixels(rc: np.ndarray, pitch_deg: float, w: int, h: int) -> np.ndarray:
    rc = np.ascontiguousarray(rc, np.int32).reshape(-1, 2)
    out = np.empty_like(rc)
    ref().ref_rotate_pixels(_p(rc, C.c_int), len(rc), pitch_deg, w, h, _p(out, C.c_int))
    return out


def ref_crop_rotated_image(im: np.ndarray, pitch_deg: float, nthreads: int = 0) -> np.ndarray:
    im = np.ascontiguousarray(im, np.uint8)
    h, w, _ = im.shape
    out = np.empty((h // 4, w, 3), np.uint8)
    ref().ref_crop_rotated_image(_p(im, C.c_uint8), w, h, pitch_deg, nthreads or max_threads(), _p(out, C.c_uint8))
    return out


def ref_rotate_keypoints(xy: np.ndarray, pitch_inv_deg: float, w: int, h: int) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2).copy()
    ref().ref_rotate_keypoints(pitch_inv_deg, _p(xy, C.c_float), len(xy), w, h)
    return xy
