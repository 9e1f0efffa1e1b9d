"""Python host layer over the C ABI (tests, bench and scripting; C++ users take host/*.hpp).

Every array argument may be a NumPy array (host memory: the library stages it, runs on the GPU and
copies results back) or a CUDA ``torch.Tensor`` (device memory: zero copy, asynchronous on the
tensor's current stream).  Results come back in the same kind.  The compute is always the CUDA
library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import MATCH_AUTO, MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16, SBA_MEM_DEVICE, SBA_MEM_HOST, SbaError, check

try:  # torch is only plumbing: device memory, streams, torch.distributed
    import torch
except Exception:  # pragma: no cover
    torch = None

_CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy handle: torch's default stream has the raw handle 0


def _is_tensor(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor)


def _ptr(a):
    if a is None:
        return None
    if _is_tensor(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _mem_of(*arrs) -> int:
    kinds = {("dev" if (_is_tensor(a) and a.is_cuda) else "host") for a in arrs if a is not None}
    if len(kinds) > 1:
        raise SbaError("mixing host and device arrays in one call")
    return SBA_MEM_DEVICE if kinds == {"dev"} else SBA_MEM_HOST


def _as(a, dtype_np, dtype_t):
    """Contiguous array of the right dtype, NumPy or torch."""
    if _is_tensor(a):
        if a.dtype != dtype_t:
            a = a.to(dtype_t)
        return a.contiguous()
    return np.ascontiguousarray(a, dtype_np)


def _empty_like_kind(ref, shape, dtype_np, dtype_t):
    if _is_tensor(ref):
        return torch.empty(shape, dtype=dtype_t, device=ref.device)
    return np.empty(shape, dtype_np)


@dataclass
class MatchResult:
    query_idx: object
    train_idx: object
    distance: object
    knn_idx: object = None
    knn_dist: object = None

    def __len__(self):
        return int(self.query_idx.shape[0])


class Context:
    """One CUDA context of the library bound to one device (wraps ``sba_ctx``)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        if stream is None and torch is not None and torch.cuda.is_available():
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream().cuda_stream or _CUDA_STREAM_LEGACY
        check(self._lib.sba_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self._problems = weakref.WeakSet()

    def close(self):
        if self._h:
            for p in list(getattr(self, "_problems", ())):  # problems hold a pointer to this context
                p.close()
            self._lib.sba_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- plumbing
    def use_current_torch_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream or _CUDA_STREAM_LEGACY
        check(self._lib.sba_ctx_set_stream(self._h, C.c_void_p(s)))

    def set_stream(self, handle: int):
        check(self._lib.sba_ctx_set_stream(self._h, C.c_void_p(handle or _CUDA_STREAM_LEGACY)))

    def synchronize(self):
        check(self._lib.sba_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.sba_ctx_launch_count(self._h))

    def set_matcher_ctas(self, n_ctas: int):
        """Persistent CTAs of the tensor-core matcher (0 = one per SM)."""
        check(self._lib.sba_ctx_set_matcher_ctas(self._h, int(n_ctas)))

    def set_dependent_launch(self, enable: bool):
        """Programmatic dependent launch along the kernel chain of a match / a pair: lower one-pair latency, ~3 % less throughput with
        several pairs in flight (default off)."""
        check(self._lib.sba_ctx_set_dependent_launch(self._h, int(bool(enable))))

    def set_remap_kernel(self, mode: int):
        """0 = per-plan choice from the timed trial, 1 = direct gather, 2 = tiled / 3 = source-ordered gather wherever possible."""
        check(self._lib.sba_ctx_set_remap_kernel(self._h, int(mode)))

    def remap_plan_info(self, w: int, h: int, cube_size: int) -> dict:
        a, p, n, f = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        ms = (C.c_float * 4)()
        check(self._lib.sba_remap_plan_info(self._h, w, h, cube_size, C.byref(a), C.byref(p), C.byref(n), C.byref(f), C.byref(ms)))
        return dict(tiled_available=bool(a.value), tiled_preferred_small=bool(p.value & 1), tiled_preferred_large=bool(p.value & 2),
                    n_tiles=n.value, n_fallback_tiles=f.value, trial_ms=dict(small_direct=float(ms[0]), small_tiled=float(ms[1]),
                                                                             large_direct=float(ms[2]), large_tiled=float(ms[3])),
                    sorted=self.remap_plan_sorted_info(w, h, cube_size))

    def remap_plan_sorted_info(self, w: int, h: int, cube_size: int) -> dict:
        """The source-ordered form of a cube plan and which kernel (1 direct, 2 tiled, 3 source-ordered) the trials picked."""
        a, bs, bl, n, f = (C.c_int32(0) for _ in range(5))
        ms = (C.c_float * 2)()
        check(self._lib.sba_remap_plan_sorted_info(self._h, w, h, cube_size, C.byref(a), C.byref(bs), C.byref(bl), C.byref(n), C.byref(f),
                                                   C.byref(ms)))
        return dict(available=bool(a.value), best_small=bs.value, best_large=bl.value, n_tiles=n.value, n_fallback_tiles=f.value,
                    trial_ms=dict(small=float(ms[0]), large=float(ms[1])))

    def set_profiling(self, enable: bool):
        check(self._lib.sba_ctx_set_profiling(self._h, int(enable)))

    def kernel_ms(self, kernel_id: int) -> float:
        """Device time of the last profiled launch: 0 matcher, 1 remap gather, 2 BA evaluation."""
        ms = C.c_float(0)
        check(self._lib.sba_ctx_kernel_ms(self._h, kernel_id, C.byref(ms)))
        return float(ms.value)

    # -- equi2cube.hpp:20-32
    def equi2cube(self, erp, cube_size: int, out=None):
        """``equi2cube::get_all`` for one image [h,w,3] or a batch [n,h,w,3] (uint8, BGR)."""
        erp = _as(erp, np.uint8, torch.uint8 if torch else None)
        batched = erp.ndim == 4
        n = erp.shape[0] if batched else 1
        h, w = (erp.shape[1], erp.shape[2]) if batched else (erp.shape[0], erp.shape[1])
        shape = (n, cube_size, 6 * cube_size, 3) if batched else (cube_size, 6 * cube_size, 3)
        if out is None:
            out = _empty_like_kind(erp, shape, np.uint8, torch.uint8 if torch else None)
        check(self._lib.sba_equi2cube(self._h, _ptr(erp), w, h, n, cube_size, _ptr(out), _mem_of(erp, out)))
        return out

    def equi2cube_face(self, erp, cube_size: int, face: int):
        """``equi2cube::get_left/front/right/back/top/bottom`` (face 0..5 in strip order)."""
        erp = _as(erp, np.uint8, torch.uint8 if torch else None)
        h, w = erp.shape[0], erp.shape[1]
        out = _empty_like_kind(erp, (cube_size, cube_size, 3), np.uint8, torch.uint8 if torch else None)
        check(self._lib.sba_equi2cube_face(self._h, _ptr(erp), w, h, cube_size, face, _ptr(out), _mem_of(erp, out)))
        return out

    def equi2cube_lut(self, w: int, h: int, cube_size: int) -> np.ndarray:
        lut = np.empty((cube_size, 6 * cube_size), np.int32)
        check(self._lib.sba_equi2cube_lut(self._h, w, h, cube_size, _ptr(lut), SBA_MEM_HOST))
        return lut

    # -- spherical_bundle_adjuster.cpp:47-181 (eight_point_estimation, initial_guess)
    def eight_point_null(self, b1, b2, idx):
        """Per subset (rows of idx [n_samples, sample_n]): null direction e [n_samples, 9] and packed A^T A [n_samples, 45]."""
        b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
        b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, np.int32)
        ns, sn = idx.shape
        e, ata = np.empty((ns, 9)), np.empty((ns, 45))
        check(self._lib.sba_eight_point_null(self._h, _ptr(b1), _ptr(b2), len(b1), _ptr(idx), ns, sn, _ptr(ata), _ptr(e), SBA_MEM_HOST))
        return e, ata

    def essential_to_candidates(self, e):
        e = np.ascontiguousarray(e, np.float64).reshape(9)
        R1, R2, T = np.empty(3, np.float32), np.empty(3, np.float32), np.empty(3, np.float32)
        v1, v2 = C.c_int32(0), C.c_int32(0)
        check(self._lib.sba_essential_to_candidates(_ptr(e), _ptr(R1), _ptr(R2), _ptr(T), C.byref(v1), C.byref(v2)))
        return R1, R2, T, bool(v1.value), bool(v2.value)

    def initial_guess(self, b1, b2, idx):
        """``spherical_bundle_adjuster::initial_guess`` with explicit subsets: (R_vec [3] f32, T_vec [3] f32, n_candidates)."""
        b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
        b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, np.int32)
        ns, sn = idx.shape
        R, T, nc = np.empty(3, np.float32), np.empty(3, np.float32), C.c_int32(0)
        check(self._lib.sba_initial_guess(self._h, _ptr(b1), _ptr(b2), len(b1), _ptr(idx), ns, sn, _ptr(R), _ptr(T), C.byref(nc), SBA_MEM_HOST))
        return R, T, int(nc.value)

    # -- spherical_surf.hpp:16-20 (eular2rot, rotate_pixel, crop_rotated_image, rotate_keypoint)
    def eular2rot(self, theta) -> np.ndarray:
        th = np.ascontiguousarray(theta, np.float32).reshape(3)
        R = np.empty(9, np.float64)
        check(self._lib.sba_eular2rot(_ptr(th), _ptr(R)))
        return R.reshape(3, 3)

    def crop_rotated_lut(self, w: int, h: int, pitch_deg: float):
        """Source-index table of one pitched band ((h/4) x w, -1 = no source) and how many entries the host settled."""
        lut = np.empty((h // 4, w), np.int32)
        n = C.c_int32(0)
        check(self._lib.sba_crop_rotated_lut(self._h, w, h, float(pitch_deg), _ptr(lut), C.byref(n), SBA_MEM_HOST))
        return lut, int(n.value)

    def crop_rotated_image(self, erp, pitch_deg: float):
        """``spherical_surf::crop_rotated_image``; erp [h, w, 3] or [n, h, w, 3] -> [.., h/4, w, 3]."""
        erp = _as(erp, np.uint8, torch.uint8 if torch else None)
        batched = erp.ndim == 4
        n = erp.shape[0] if batched else 1
        h, w = erp.shape[-3], erp.shape[-2]
        shape = (n, h // 4, w, 3) if batched else (h // 4, w, 3)
        out = _empty_like_kind(erp, shape, np.uint8, torch.uint8 if torch else None)
        check(self._lib.sba_crop_rotated_image(self._h, _ptr(erp), w, h, n, float(pitch_deg), _ptr(out), _mem_of(erp, out)))
        return out

    def spherical_crops(self, erp):
        """The four bands of ``spherical_surf::do_all`` (pitch 45, plain band, -45, -90): [.., 4, h/4, w, 3]."""
        erp = _as(erp, np.uint8, torch.uint8 if torch else None)
        batched = erp.ndim == 4
        n = erp.shape[0] if batched else 1
        h, w = erp.shape[-3], erp.shape[-2]
        shape = (n, 4, h // 4, w, 3) if batched else (4, h // 4, w, 3)
        out = _empty_like_kind(erp, shape, np.uint8, torch.uint8 if torch else None)
        check(self._lib.sba_spherical_crops(self._h, _ptr(erp), w, h, n, _ptr(out), _mem_of(erp, out)))
        return out

    def rotate_pixels(self, rc, pitch_deg: float, w: int, h: int):
        rc = _as(rc, np.int32, torch.int32 if torch else None).reshape(-1, 2)
        out = _empty_like_kind(rc, rc.shape, np.int32, torch.int32 if torch else None)
        check(self._lib.sba_rotate_pixels(self._h, _ptr(rc), rc.shape[0], float(pitch_deg), w, h, _ptr(out), _mem_of(rc, out)))
        return out

    def rotate_pixels_mat(self, rc, R, w: int, h: int):
        """``spherical_surf::rotate_pixel`` with an arbitrary rotation matrix (3x3 doubles)."""
        rc = np.ascontiguousarray(rc, np.int32).reshape(-1, 2)
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        out = np.empty_like(rc)
        check(self._lib.sba_rotate_pixels_mat(self._h, _ptr(rc), rc.shape[0], _ptr(R), w, h, _ptr(out), SBA_MEM_HOST))
        return out

    def rotate_keypoints(self, xy, pitch_inv_deg: float, w: int, h: int):
        """``spherical_surf::rotate_keypoint`` on a copy of xy [n, 2]."""
        xy = _as(xy, np.float32, torch.float32 if torch else None).reshape(-1, 2)
        xy = xy.clone() if _is_tensor(xy) else xy.copy()
        check(self._lib.sba_rotate_keypoints(self._h, _ptr(xy), xy.shape[0], float(pitch_inv_deg), w, h, _mem_of(xy)))
        return xy

    # -- equi2cube_surf.hpp:12
    def cube2equi_points(self, xy, cube_size: int, w: int, h: int):
        xy = _as(xy, np.float32, torch.float32 if torch else None).reshape(-1, 2)
        out = _empty_like_kind(xy, xy.shape, np.float32, torch.float32 if torch else None)
        check(self._lib.sba_cube2equi_points(self._h, _ptr(xy), xy.shape[0], cube_size, w, h, _ptr(out), _mem_of(xy, out)))
        return out

    # -- spherical_bundle_adjuster.cpp:271-298
    def pixels_to_bearings(self, xy, w: int, h: int, want_f64: bool = False):
        xy = _as(xy, np.float32, torch.float32 if torch else None).reshape(-1, 2)
        n = xy.shape[0]
        b32 = _empty_like_kind(xy, (n, 4), np.float32, torch.float32 if torch else None)
        b64 = _empty_like_kind(xy, (n, 3), np.float64, torch.float64 if torch else None) if want_f64 else None
        check(self._lib.sba_pixels_to_bearings(self._h, _ptr(xy), n, w, h, _ptr(b32), _ptr(b64), _mem_of(xy, b32)))
        return (b32, b64) if want_f64 else b32

    # -- feature_matcher.hpp:34
    def match_two_image(self, desc1, desc2, ratio: float = 0.3, algo: int = MATCH_AUTO, want_knn: bool = False) -> MatchResult:
        """``feature_matcher::match_two_image``: kNN(k=2) + ratio test, survivors in query order."""
        if isinstance(desc1, Descriptors) and isinstance(desc2, Descriptors):
            nq = len(desc1)
            qi, ti, dd = np.empty(max(nq, 1), np.int32), np.empty(max(nq, 1), np.int32), np.empty(max(nq, 1), np.float32)
            ki = np.empty((nq, 2), np.int32) if want_knn else None
            kd = np.empty((nq, 2), np.float32) if want_knn else None
            nm = C.c_int32(0)
            check(self._lib.sba_knn2_ratio_prepared(self._h, desc1._h, desc2._h, ratio, _ptr(qi), _ptr(ti), _ptr(dd), C.cast(C.byref(nm), C.c_void_p),
                                                    _ptr(ki), _ptr(kd), SBA_MEM_HOST, algo))
            return MatchResult(qi[:nm.value], ti[:nm.value], dd[:nm.value], ki, kd)
        f32t = torch.float32 if torch else None
        i32t = torch.int32 if torch else None
        q = _as(desc1, np.float32, f32t)
        t = _as(desc2, np.float32, f32t)
        nq = q.shape[0]
        nt = t.shape[0]
        dim = q.shape[1] if q.ndim == 2 and nq else (t.shape[1] if t.ndim == 2 else 64)
        qi = _empty_like_kind(q, (max(nq, 1),), np.int32, i32t)
        ti = _empty_like_kind(q, (max(nq, 1),), np.int32, i32t)
        dd = _empty_like_kind(q, (max(nq, 1),), np.float32, f32t)
        ki = _empty_like_kind(q, (nq, 2), np.int32, i32t) if want_knn else None
        kd = _empty_like_kind(q, (nq, 2), np.float32, f32t) if want_knn else None
        mem = _mem_of(q, t)
        if mem == SBA_MEM_DEVICE:
            nm = torch.zeros(1, dtype=torch.int32, device=q.device)
            check(self._lib.sba_knn2_ratio(self._h, _ptr(q), nq, _ptr(t), nt, dim, ratio, _ptr(qi), _ptr(ti), _ptr(dd), _ptr(nm),
                                           _ptr(ki), _ptr(kd), mem, algo))
            n = int(nm.item())
        else:
            nm = C.c_int32(0)
            check(self._lib.sba_knn2_ratio(self._h, _ptr(q), nq, _ptr(t), nt, dim, ratio, _ptr(qi), _ptr(ti), _ptr(dd),
                                           C.cast(C.byref(nm), C.c_void_p), _ptr(ki), _ptr(kd), mem, algo))
            n = nm.value
        return MatchResult(qi[:n], ti[:n], dd[:n], ki, kd)

    def match_stats(self) -> _lib.MatchStats:
        s = _lib.MatchStats()
        check(self._lib.sba_match_last_stats(self._h, C.byref(s)))
        return s

    def prepare_descriptors(self, desc) -> "Descriptors":
        """Hand a descriptor set over once (``sba_descriptors_create``); pass the result to match_two_image / match_begin
        in place of the array for every match it takes part in."""
        return Descriptors(self, desc)

    def match_begin(self, desc1, desc2, ratio: float = 0.3, algo: int = MATCH_AUTO) -> "MatchCall":
        """Queue ``match_two_image`` on CUDA tensors and return at once (nothing is copied to the host);
        ``.end()`` waits for this context's stream and returns the MatchResult.  Several contexts on their own
        streams keep several image pairs in flight (all-pairs matching of a sequence, BASELINE config 3)."""
        prepared = isinstance(desc1, Descriptors) and isinstance(desc2, Descriptors)
        if prepared:
            q, t, nq = desc1, desc2, len(desc1)
            dev = torch.device("cuda", self.device)
        else:
            q = _as(desc1, np.float32, torch.float32)
            t = _as(desc2, np.float32, torch.float32)
            if _mem_of(q, t) != SBA_MEM_DEVICE:
                raise SbaError("match_begin takes CUDA tensors or prepared descriptor sets")
            nq, nt, dim, dev = q.shape[0], t.shape[0], q.shape[1], q.device
        qi = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)
        ti = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)
        dd = torch.empty(max(nq, 1), dtype=torch.float32, device=dev)
        nm = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.current_stream(dev).synchronize()      # the count was just zeroed on torch's stream
        if prepared:
            check(self._lib.sba_knn2_ratio_prepared(self._h, q._h, t._h, ratio, _ptr(qi), _ptr(ti), _ptr(dd), _ptr(nm), None, None, SBA_MEM_DEVICE, algo))
        else:
            check(self._lib.sba_knn2_ratio(self._h, _ptr(q), nq, _ptr(t), nt, dim, ratio, _ptr(qi), _ptr(ti), _ptr(dd), _ptr(nm), None, None,
                                           SBA_MEM_DEVICE, algo))
        return MatchCall(self, (q, t), qi, ti, dd, nm)

    def gather_matches(self, key_left_xy, key_right_xy, query_idx, train_idx):
        n = int(query_idx.shape[0])
        ol = torch.empty((n, 2), dtype=torch.float32, device=key_left_xy.device)
        orr = torch.empty((n, 2), dtype=torch.float32, device=key_left_xy.device)
        check(self._lib.sba_gather_matches(self._h, _ptr(key_left_xy.contiguous()), _ptr(key_right_xy.contiguous()),
                                           _ptr(query_idx.contiguous()), _ptr(train_idx.contiguous()), n, _ptr(ol), _ptr(orr),
                                           SBA_MEM_DEVICE))
        return ol, orr

    # -- the whole hot path for one pair (equi2cube_surf::do_all + rotation-only BA) in one C call
    def pair_rotation(self, im_left, im_right, desc_left, desc_right, key_left_xy, key_right_xy, cube_size: int, w: int = 0, h: int = 0,
                      ratio: float = 0.3, r0=(0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0), d1=1.0, d2=1.0, huber=1.0, max_iter=50,
                      want_matches: bool = True, want_strips: bool = False):
        """Returns (PairResult, matches (query_idx, train_idx, dist) or None, strips or None)."""
        return self.pair_rotation_begin(im_left, im_right, desc_left, desc_right, key_left_xy, key_right_xy, cube_size, w, h, ratio, r0, t,
                                        d1, d2, huber, max_iter, want_matches, want_strips, _blocking=True).end()

    def pair_rotation_begin(self, im_left, im_right, desc_left, desc_right, key_left_xy, key_right_xy, cube_size: int, w: int = 0, h: int = 0,
                            ratio: float = 0.3, r0=(0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0), d1=1.0, d2=1.0, huber=1.0, max_iter=50,
                            want_matches: bool = True, want_strips: bool = False, _blocking: bool = False, out: dict | None = None) -> "PairCall":
        """Queue one pair and return at once; ``.end()`` on the returned call collects what ``pair_rotation`` returns.
        One call in flight per context: keep N pairs in flight with N contexts on N streams.
        ``out`` may hold preallocated result buffers (keys qi, ti, dd, sl, sr) of the same kind as the inputs --
        pinned host tensors keep the device -> host copies asynchronous."""
        u8t = torch.uint8 if torch else None
        f32t = torch.float32 if torch else None
        i32t = torch.int32 if torch else None
        dl, dr = _as(desc_left, np.float32, f32t), _as(desc_right, np.float32, f32t)
        kl, kr = _as(key_left_xy, np.float32, f32t), _as(key_right_xy, np.float32, f32t)
        iml = imr = None
        if im_left is not None:
            iml, imr = _as(im_left, np.uint8, u8t), _as(im_right, np.uint8, u8t)
            h, w = iml.shape[0], iml.shape[1]
        mem = _mem_of(iml, imr, dl, dr, kl, kr)
        nl, nr = dl.shape[0], dr.shape[0]
        dim = dl.shape[1]
        out = out or {}
        qi = out.get("qi", _empty_like_kind(dl, (max(nl, 1),), np.int32, i32t)) if want_matches else None
        ti = out.get("ti", _empty_like_kind(dl, (max(nl, 1),), np.int32, i32t)) if want_matches else None
        dd = out.get("dd", _empty_like_kind(dl, (max(nl, 1),), np.float32, f32t)) if want_matches else None
        sl = sr = None
        if want_strips and iml is not None:
            sl = out.get("sl", _empty_like_kind(dl, (cube_size, 6 * cube_size, 3), np.uint8, u8t))
            sr = out.get("sr", _empty_like_kind(dl, (cube_size, 6 * cube_size, 3), np.uint8, u8t))
        r0 = np.ascontiguousarray(r0, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        call = PairCall(self, (iml, imr, dl, dr, kl, kr, r0, t), qi, ti, dd, sl, sr)
        args = (self._h, _ptr(iml), _ptr(imr), w, h, cube_size, _ptr(sl), _ptr(sr), _ptr(dl), nl, _ptr(dr), nr, dim,
                _ptr(kl), _ptr(kr), ratio, _ptr(r0), _ptr(t), d1, d2, huber, max_iter, _ptr(qi), _ptr(ti), _ptr(dd))
        if _blocking:   # the one-call entry point
            check(self._lib.sba_pair_rotation(*args, C.byref(call._res), mem))
            call._done = True
        else:
            check(self._lib.sba_pair_rotation_begin(*args, mem, C.byref(call._h)))
        return call

    # -- bundle adjustment
    def ba_problem(self, b1, b2, cam=None, n_cam: int = 1) -> "BAProblem":
        return BAProblem(self, b1, b2, cam, n_cam)


class Descriptors:
    """A descriptor set resident on the device in the matcher's prepared form (wraps ``sba_descriptors``)."""

    def __init__(self, ctx: "Context", desc):
        self._lib = ctx._lib
        self._h = C.c_void_p()
        d = _as(desc, np.float32, torch.float32 if torch else None)
        check(self._lib.sba_descriptors_create(ctx._h, _ptr(d), d.shape[0], d.shape[1] if d.ndim == 2 else 64, _mem_of(d), C.byref(self._h)))

    def __len__(self):
        return int(self._lib.sba_descriptors_count(self._h))

    def close(self):
        if self._h:
            self._lib.sba_descriptors_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MatchCall:
    """A match queued by ``Context.match_begin``; keeps its buffers alive until ``end``."""

    def __init__(self, ctx, inputs, qi, ti, dd, nm):
        self._ctx, self._inputs, self._qi, self._ti, self._dd, self._nm = ctx, inputs, qi, ti, dd, nm

    def end(self) -> "MatchResult":
        self._ctx.synchronize()
        n = int(self._nm.item())
        return MatchResult(self._qi[:n], self._ti[:n], self._dd[:n], None, None)


class PairCall:
    """A pair queued by ``Context.pair_rotation_begin`` (wraps ``sba_pair_call``); keeps every buffer alive until ``end``."""

    def __init__(self, ctx, inputs, qi, ti, dd, sl, sr):
        self._ctx, self._inputs = ctx, inputs
        self._qi, self._ti, self._dd, self._sl, self._sr = qi, ti, dd, sl, sr
        self._h = C.c_void_p()
        self._res = _lib.PairResult()
        self._done = False

    def end(self):
        if not self._done:
            check(self._ctx._lib.sba_pair_rotation_end(self._h, C.byref(self._res)))
            self._done = True
            self._h = C.c_void_p()
        res, n = self._res, self._res.n_matches
        matches = (self._qi[:n], self._ti[:n], self._dd[:n]) if self._qi is not None else None
        return res, matches, ((self._sl, self._sr) if self._sl is not None else None)

    def __del__(self):
        try:
            if not self._done and self._h:
                self.end()
        except Exception:
            pass


class PeerComm:
    """Peer-memory exchange group for residual-sharded BA (wraps ``sba_comm``): one per process, all GPUs of
    one NVLink box.  ``torch.distributed`` only carries the 64-byte IPC handles at set-up."""

    def __init__(self, ctx: "Context", rank: int, world: int, max_cameras: int, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        check(self._lib.sba_comm_create(ctx._h, rank, world, max_cameras, C.byref(self._h), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self._lib.sba_comm_connect(self._h, buf))
        dist.barrier(group=group)

    def close(self):
        if self._h:
            self._lib.sba_comm_destroy(self._h)
            self._h = C.c_void_p()


def _bearings4(b, f32t):
    """n x 3 or n x 4 -> contiguous n x 4 float32 (x, y, z, 0)."""
    if _is_tensor(b):
        b = b.to(torch.float32)
        if b.shape[1] == 3:
            b = torch.cat([b, torch.zeros_like(b[:, :1])], dim=1)
        return b.contiguous()
    b = np.asarray(b, np.float32)
    if b.shape[1] == 3:
        b = np.concatenate([b, np.zeros_like(b[:, :1])], axis=1)
    return np.ascontiguousarray(b)


class BAProblem:
    """Rotation-only BA problem resident on the device (wraps ``sba_ba_problem``)."""

    def __init__(self, ctx: Context, b1, b2, cam=None, n_cam: int = 1):
        f32t = torch.float32 if torch else None
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = C.c_void_p()
        b1 = _bearings4(b1, f32t)
        b2 = _bearings4(b2, f32t)
        if cam is not None:
            cam = _as(cam, np.int32, torch.int32 if torch else None)
        self.n_obs = int(b1.shape[0])
        self.n_cam = int(n_cam)
        self._device_kind = _is_tensor(b1) and b1.is_cuda
        check(self._lib.sba_ba_problem_create(ctx._h, _ptr(b1), _ptr(b2), _ptr(cam), self.n_obs, self.n_cam, _mem_of(b1, b2, cam),
                                              C.byref(self._h)))
        self._cb = None
        ctx._problems.add(self)

    def close(self):
        if self._h and self.ctx._h:
            self._lib.sba_ba_problem_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_allreduce(self, fn):
        """``fn(ptr:int, count:int) -> None`` sums `count` doubles at device address `ptr` over ranks."""
        if fn is None:
            self._cb = None
            check(self._lib.sba_ba_problem_set_allreduce(self._h, _lib.ALLREDUCE_FN(), None))
            return

        def _tramp(ptr, count, _user):
            try:
                fn(ptr, count)
                return 0
            except Exception as e:  # pragma: no cover
                print("allreduce callback failed:", e)
                return 1

        self._cb = _lib.ALLREDUCE_FN(_tramp)
        check(self._lib.sba_ba_problem_set_allreduce(self._h, self._cb, None))

    def set_comm(self, comm: "PeerComm | None"):
        """Sum the per-camera blocks over ranks inside the evaluation kernel (NVLink peer memory)."""
        self._comm = comm   # keep it alive
        check(self._lib.sba_ba_problem_set_comm(self._h, comm._h if comm is not None else None))

    @staticmethod
    def _r(r, n_cam):
        r = np.ascontiguousarray(np.asarray(r, np.float64).reshape(n_cam, 3))
        return r

    def eval(self, r, t=(0.0, 0.0, 0.0), d1=1.0, d2=1.0, huber=1.0, want_res=False, want_jac=False, device_out=False):
        """One evaluation.  Returns dict(res, jac, H, g, cost); NumPy unless device_out."""
        r = self._r(r, self.n_cam)
        t = np.ascontiguousarray(t, np.float64)
        n, m = self.n_obs, self.n_cam
        if device_out:
            dev = torch.device("cuda", self.ctx.device)
            res = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_res else None
            jac = torch.empty((n, 3, 3), dtype=torch.float32, device=dev) if want_jac else None
            H = torch.empty((m, 6), dtype=torch.float64, device=dev)
            g = torch.empty((m, 3), dtype=torch.float64, device=dev)
            cost = torch.empty((m,), dtype=torch.float64, device=dev)
            mem = SBA_MEM_DEVICE
        else:
            res = np.empty((n, 3), np.float32) if want_res else None
            jac = np.empty((n, 3, 3), np.float32) if want_jac else None
            H = np.empty((m, 6)); g = np.empty((m, 3)); cost = np.empty(m)
            mem = SBA_MEM_HOST
        check(self._lib.sba_ba_rot_eval(self._h, _ptr(r), _ptr(t), d1, d2, huber, _ptr(res), _ptr(jac), _ptr(H), _ptr(g), _ptr(cost), mem))
        return dict(res=res, jac=jac, H=H, g=g, cost=cost)

    def solve(self, r0, t=(0.0, 0.0, 0.0), d1=1.0, d2=1.0, huber=1.0, max_iter=50):
        r = self._r(r0, self.n_cam).copy()
        t = np.ascontiguousarray(t, np.float64)
        s = _lib.SolveSummary()
        check(self._lib.sba_ba_rot_solve(self._h, _ptr(r), _ptr(t), d1, d2, huber, max_iter, C.byref(s)))
        return r, s

    # -- translation-only block (ba_spherical_costfunctor_tran_only, spherical_bundle_adjuster.cpp:948-1002)
    def tran_eval(self, r_fixed, t, d1=1.0, d2=1.0, huber=1.0, want_res=False):
        r = self._r(r_fixed, self.n_cam)
        tv = self._r(t, self.n_cam)
        n, m = self.n_obs, self.n_cam
        res = np.empty((n, 3), np.float32) if want_res else None
        H = np.empty((m, 6)); g = np.empty((m, 3)); cost = np.empty(m)
        check(self._lib.sba_ba_tran_eval(self._h, _ptr(r), _ptr(tv), d1, d2, huber, _ptr(res), _ptr(H), _ptr(g), _ptr(cost), SBA_MEM_HOST))
        return dict(res=res, H=H, g=g, cost=cost)

    def tran_solve(self, r_fixed, t0, d1=1.0, d2=1.0, huber=1.0, max_iter=50):
        r = self._r(r_fixed, self.n_cam)
        tv = self._r(t0, self.n_cam).copy()
        s = _lib.SolveSummary()
        check(self._lib.sba_ba_tran_solve(self._h, _ptr(r), _ptr(tv), d1, d2, huber, max_iter, C.byref(s)))
        return tv, s

    # ---- depth-only block and the three-stage sequence (single camera pair) --------------------------------
    def _d(self, d):
        d = np.ascontiguousarray(d, np.float64).reshape(-1, 2)
        if len(d) != self.n_obs:
            raise ValueError(f"expected {self.n_obs} depth pairs, got {len(d)}")
        return d

    def d_eval(self, r, t, d, lam=1.0, c=1.0):
        """Raw depth-only functor values (res [n, 5], jac [n, 5, 2], cost [n]); spherical_bundle_adjuster.cpp:1005-1032."""
        r, t, d = self._r(r, 1).reshape(3), np.ascontiguousarray(t, np.float64).reshape(3), self._d(d)
        res, jac, cost = np.empty((self.n_obs, 5)), np.empty((self.n_obs, 5, 2)), np.empty(self.n_obs)
        check(self._lib.sba_ba_d_eval(self._h, _ptr(r), _ptr(t), _ptr(d), lam, c, _ptr(res), _ptr(jac), _ptr(cost), SBA_MEM_HOST))
        return res, jac, cost

    def d_solve(self, r, t, d0, lam=1.0, c=1.0, max_iter=50):
        """Depth stage of solve_problem (:196-197): returns (d [n, 2], summary, line-search trials beyond alpha = 1)."""
        r, t = self._r(r, 1).reshape(3), np.ascontiguousarray(t, np.float64).reshape(3)
        if _is_tensor(d0) and d0.is_cuda:   # device mode: depths stay in HBM
            d = d0.to(torch.float64).reshape(-1, 2).contiguous().clone()
            if d.shape[0] != self.n_obs:
                raise ValueError(f"expected {self.n_obs} depth pairs, got {d.shape[0]}")
            mem = SBA_MEM_DEVICE
        else:
            d, mem = self._d(d0).copy(), SBA_MEM_HOST
        s, nls = _lib.SolveSummary(), C.c_int32(0)
        check(self._lib.sba_ba_d_solve(self._h, _ptr(r), _ptr(t), _ptr(d), lam, c, max_iter, C.byref(s), C.byref(nls), mem))
        return d, s, int(nls.value)

    def solve_problem(self, r0, t0, d0, huber=1.0, max_iter=50):
        """spherical_bundle_adjuster::solve_problem (:183-217): depth -> rotation -> translation.
        Returns (r [3], t [3], d [n, 2], [summary_d, summary_rot, summary_tran]); d0 may be a CUDA tensor (stays on the device)."""
        r, t = self._r(r0, 1).reshape(3).copy(), np.ascontiguousarray(t0, np.float64).reshape(3).copy()
        if _is_tensor(d0) and d0.is_cuda:
            d = d0.to(torch.float64).reshape(-1, 2).contiguous().clone()
            if d.shape[0] != self.n_obs:
                raise ValueError(f"expected {self.n_obs} depth pairs, got {d.shape[0]}")
            mem = SBA_MEM_DEVICE
        else:
            d, mem = self._d(d0).copy(), SBA_MEM_HOST
        sums = (_lib.SolveSummary * 3)()
        check(self._lib.sba_ba_solve_problem(self._h, _ptr(r), _ptr(t), _ptr(d), huber, max_iter, C.byref(sums), mem))
        return r, t, d, list(sums)

    def eval_timed(self, r, t=(0.0, 0.0, 0.0), d1=1.0, d2=1.0, huber=1.0, materialise=False, iters=20) -> float:
        r = self._r(r, self.n_cam)
        t = np.ascontiguousarray(t, np.float64)
        ms = C.c_float(0)
        check(self._lib.sba_ba_rot_eval_timed(self._h, _ptr(r), _ptr(t), d1, d2, huber, int(materialise), iters, C.byref(ms)))
        return float(ms.value)
