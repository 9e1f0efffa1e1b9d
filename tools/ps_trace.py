"""Timeline of ba_pair_solve_kernel (debug build: make -C spherical_bundle_adjuster_b200/csrc trace; SBA_B200_LIB=build/libsba_b200_trace.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, synth, _lib
W, H, CS, N = 3840, 1920, 960, 16384
ctx = Context(0)
pair = synth.make_pair(N, N, cs=CS, seed=5)
d = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (synth.make_erp_image(W, H, 1), synth.make_erp_image(W, H, 2), pair["desc1"], pair["desc2"], pair["key1_xy"], pair["key2_xy"])]
for _ in range(4):
    res, m, _ = ctx.pair_rotation(*d, CS)
buf = np.zeros(64, np.uint64)
assert _lib.load().sba_ps_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
t = (buf.astype(np.int64) - int(buf[0])) / 1000.0
print("start 0; all CTAs up", t[1], "; bearings done", t[2])
for g in range(res.lm_iterations + 1):
    print(f"gen {g}: eval done {t[3+4*g]:.2f}, barrier {t[4+4*g]:.2f}, decide done {t[5+4*g]:.2f}, barrier {t[6+4*g]:.2f} us")
