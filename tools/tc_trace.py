"""Per-CTA timeline of the tensor-core matcher kernel (debug build: `make -C spherical_bundle_adjuster_b200/csrc trace`,
run with SBA_B200_LIB=build/libsba_b200_trace.so).  Prints where the kernel's time goes that is not MMA work."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherical_bundle_adjuster_b200 import Context, synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = Context(0)
A, B, _ = synth.make_descriptors(n, n, 64, seed=1)
a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
lib = _lib.load()
for _ in range(5):
    m = ctx.match_two_image(a, b, 0.3)
torch.cuda.synchronize()
edges = np.zeros(2, np.uint64)
if hasattr(lib, "sba_tc_trace_edges"):
    lib.sba_tc_trace_edges(edges.ctypes.data_as(C.c_void_p), 1)
    m = ctx.match_two_image(a, b, 0.3)
    torch.cuda.synchronize()
    lib.sba_tc_trace_edges(edges.ctypes.data_as(C.c_void_p), 0)
buf = np.zeros((148, 16), np.uint64)
assert lib.sba_tc_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
t = buf[:, :8].astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0     # microseconds since the first CTA started
names = ["cta_start", "setup_done", "first_operands", "last_tile_issue", "first_acc_ready", "last_acc_ready", "epilogue_done", "cta_end"]
print(json.dumps({"n": n, "tiles_per_cta": [int(buf[:, 8].min()), int(buf[:, 8].max())]}))
for k, nm in enumerate(names):
    print(f"{nm:18s} min {rel[:, k].min():8.2f}  median {np.median(rel[:, k]):8.2f}  max {rel[:, k].max():8.2f} us")
if edges[1] > 0:
    print("prep kernel's last thread -> first CTA start: %.2f us; last CTA end -> re-rank kernel's first thread: %.2f us" %
          ((int(t0) - int(edges[0])) / 1000.0, (int(edges[1]) - int(t[:, 7].max())) / 1000.0))
dur = rel[:, 7] - rel[:, 0]
print("per-CTA duration  min %.2f median %.2f max %.2f us; kernel span %.2f us" % (dur.min(), np.median(dur), dur.max(), rel[:, 7].max()))
steady = (rel[:, 5] - rel[:, 4]) / np.maximum(buf[:, 8].astype(np.float64) - 1, 1)
print("steady-state per tile (first->last accumulator): min %.3f median %.3f max %.3f us" % (steady.min(), np.median(steady), steady.max()))
