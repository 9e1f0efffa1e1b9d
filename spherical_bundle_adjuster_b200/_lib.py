"""ctypes binding of ``libsba_b200.so`` (the C ABI declared in ``include/sba_b200.h``).

There is no fallback: if the shared library is missing, or no sm_100 device is usable, importing
works but creating a :class:`Context` raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBA_B200_LIB") or os.path.join(_HERE, "libsba_b200.so")   # override: debug builds under build/

SBA_MEM_HOST, SBA_MEM_DEVICE = 0, 1
MATCH_AUTO, MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16 = 0, 1, 2, 3

# every symbol include/sba_b200.h declares (tests check the .so exports all of them)
EXPORTED = [
    "sba_version", "sba_last_error", "sba_ctx_create", "sba_ctx_destroy", "sba_ctx_set_stream", "sba_ctx_get_stream",
    "sba_ctx_synchronize", "sba_ctx_launch_count", "sba_ctx_set_profiling", "sba_ctx_kernel_ms", "sba_equi2cube", "sba_equi2cube_face", "sba_equi2cube_lut",
    "sba_cube2equi_points", "sba_pixels_to_bearings", "sba_knn2_ratio", "sba_match_last_stats", "sba_gather_matches",
    "sba_ba_problem_create", "sba_ba_problem_destroy", "sba_ba_problem_set_allreduce", "sba_ba_rot_eval",
    "sba_ba_rot_solve", "sba_ba_rot_eval_timed", "sba_pair_rotation", "sba_ba_tran_eval", "sba_ba_tran_solve", "sba_comm_create", "sba_comm_connect", "sba_comm_destroy",
    "sba_ba_problem_set_comm", "sba_ba_d_eval", "sba_ba_d_solve", "sba_ba_solve_problem",
    "sba_eular2rot", "sba_crop_rotated_lut", "sba_crop_rotated_image", "sba_spherical_crops", "sba_rotate_pixels", "sba_rotate_pixels_mat", "sba_rotate_keypoints",
    "sba_pair_rotation_begin", "sba_pair_rotation_end", "sba_eight_point_null", "sba_essential_to_candidates", "sba_initial_guess",
    "sba_ctx_set_remap_kernel", "sba_remap_plan_info", "sba_remap_plan_sorted_info", "sba_ctx_set_matcher_ctas", "sba_ctx_set_dependent_launch",
    "sba_descriptors_create", "sba_descriptors_destroy", "sba_descriptors_count", "sba_knn2_ratio_prepared",
]


class SbaError(RuntimeError):
    pass


class MatchStats(C.Structure):
    _fields_ = [("algo_used", C.c_int), ("n_fallback_rows", C.c_int), ("n_tiles", C.c_int), ("n_ctas", C.c_int),
                ("max_rel_err", C.c_float)]


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_successful", C.c_int), ("termination", C.c_int), ("evaluations", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double)]


class PairResult(C.Structure):
    _fields_ = [("rotation", C.c_double * 3), ("n_matches", C.c_int), ("lm_iterations", C.c_int), ("lm_termination", C.c_int),
                ("reserved", C.c_int), ("initial_cost", C.c_double), ("final_cost", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p)

_lib = None


def load():
    """Load the shared library (once).  Raises SbaError with the build hint when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbaError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    lib.sba_version.restype = i32
    lib.sba_last_error.restype = C.c_char_p
    lib.sba_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.sba_ctx_destroy.argtypes = [vp]
    lib.sba_ctx_set_stream.argtypes = [vp, vp]
    lib.sba_ctx_get_stream.argtypes = [vp]
    lib.sba_ctx_get_stream.restype = vp
    lib.sba_ctx_synchronize.argtypes = [vp]
    lib.sba_ctx_launch_count.argtypes = [vp]
    lib.sba_ctx_launch_count.restype = i64
    lib.sba_ctx_set_profiling.argtypes = [vp, i32]
    lib.sba_ctx_kernel_ms.argtypes = [vp, i32, C.POINTER(f32)]
    lib.sba_equi2cube.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32]
    lib.sba_equi2cube_face.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32]
    lib.sba_equi2cube_lut.argtypes = [vp, i32, i32, i32, vp, i32]
    lib.sba_cube2equi_points.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32]
    lib.sba_pixels_to_bearings.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32]
    lib.sba_knn2_ratio.argtypes = [vp, vp, i32, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.sba_match_last_stats.argtypes = [vp, C.POINTER(MatchStats)]
    lib.sba_gather_matches.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i32]
    lib.sba_ba_problem_create.argtypes = [vp, vp, vp, vp, i64, i32, i32, C.POINTER(vp)]
    lib.sba_ba_problem_destroy.argtypes = [vp]
    lib.sba_ba_problem_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    lib.sba_ba_rot_eval.argtypes = [vp, vp, vp, f64, f64, f64, vp, vp, vp, vp, vp, i32]
    lib.sba_ba_rot_solve.argtypes = [vp, vp, vp, f64, f64, f64, i32, C.POINTER(SolveSummary)]
    lib.sba_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp), vp]
    lib.sba_comm_connect.argtypes = [vp, vp]
    lib.sba_comm_destroy.argtypes = [vp]
    lib.sba_ba_problem_set_comm.argtypes = [vp, vp]
    lib.sba_ba_tran_eval.argtypes = [vp, vp, vp, f64, f64, f64, vp, vp, vp, vp, i32]
    lib.sba_ba_tran_solve.argtypes = [vp, vp, vp, f64, f64, f64, i32, C.POINTER(SolveSummary)]
    lib.sba_eular2rot.argtypes = [vp, vp]
    lib.sba_crop_rotated_lut.argtypes = [vp, i32, i32, f32, vp, C.POINTER(i32), i32]
    lib.sba_crop_rotated_image.argtypes = [vp, vp, i32, i32, i32, f32, vp, i32]
    lib.sba_spherical_crops.argtypes = [vp, vp, i32, i32, i32, vp, i32]
    lib.sba_rotate_pixels.argtypes = [vp, vp, i32, f32, i32, i32, vp, i32]
    lib.sba_rotate_pixels_mat.argtypes = [vp, vp, i32, vp, i32, i32, vp, i32]
    lib.sba_rotate_keypoints.argtypes = [vp, vp, i32, f32, i32, i32, i32]
    lib.sba_ba_d_eval.argtypes = [vp, vp, vp, vp, f64, f64, vp, vp, vp, i32]
    lib.sba_ba_d_solve.argtypes = [vp, vp, vp, vp, f64, f64, i32, C.POINTER(SolveSummary), C.POINTER(i32), i32]
    lib.sba_ba_solve_problem.argtypes = [vp, vp, vp, vp, f64, i32, C.POINTER(SolveSummary * 3), i32]
    lib.sba_ba_rot_eval_timed.argtypes = [vp, vp, vp, f64, f64, f64, i32, i32, C.POINTER(f32)]
    lib.sba_pair_rotation.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp, f32, vp, vp, f64, f64, f64, i32,
                                      vp, vp, vp, C.POINTER(PairResult), i32]
    lib.sba_descriptors_create.argtypes = [vp, vp, i32, i32, i32, C.POINTER(vp)]
    lib.sba_descriptors_destroy.argtypes = [vp]
    lib.sba_descriptors_count.argtypes = [vp]
    lib.sba_knn2_ratio_prepared.argtypes = [vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.sba_ctx_set_remap_kernel.argtypes = [vp, i32]
    lib.sba_ctx_set_dependent_launch.argtypes = [vp, i32]
    lib.sba_ctx_set_matcher_ctas.argtypes = [vp, i32]
    lib.sba_remap_plan_info.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(f32 * 4)]
    lib.sba_remap_plan_sorted_info.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(f32 * 2)]
    lib.sba_eight_point_null.argtypes = [vp, vp, vp, i32, vp, i32, i32, vp, vp, i32]
    lib.sba_essential_to_candidates.argtypes = [vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
    lib.sba_initial_guess.argtypes = [vp, vp, vp, i32, vp, i32, i32, vp, vp, C.POINTER(i32), i32]
    lib.sba_pair_rotation_begin.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp, f32, vp, vp, f64, f64, f64, i32,
                                            vp, vp, vp, i32, C.POINTER(vp)]
    lib.sba_pair_rotation_end.argtypes = [vp, C.POINTER(PairResult)]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise SbaError(f"libsba_b200 error {status}: {load().sba_last_error().decode()}")
