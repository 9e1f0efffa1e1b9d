"""B200-native hot path of whdlgp/spherical_bundle_adjuster: equi2cube remap, brute-force SURF
descriptor matching (kNN k=2 + ratio test) and rotation-only spherical bundle adjustment, as
hand-written sm_100a CUDA kernels behind the C ABI in ``include/sba_b200.h``."""
from ._lib import MATCH_AUTO, MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16, SbaError  # noqa: F401
from .api import BAProblem, Context, Descriptors, MatchResult, PeerComm  # noqa: F401

__all__ = ["Context", "BAProblem", "Descriptors", "MatchResult", "PeerComm", "SbaError", "MATCH_AUTO", "MATCH_SIMT_EXACT", "MATCH_TENSOR", "MATCH_TENSOR_FP16"]
