"""BASELINE config 3: all-pairs matching of a synthetic ERP sequence, pairs sharded across ranks (no collective).
  python tools/bench_allpairs.py [frames] [descriptors per frame] [contexts in flight]
  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/bench_allpairs.py ...
Descriptors of every frame are resident on every GPU (64 x 16k x 256 B = 268 MB); each rank matches its contiguous share
of the pair list with a few library contexts in flight.  Prints one JSON line (rank 0): whole-job pairs/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from spherical_bundle_adjuster_b200 import Context, sharding

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
PREP = int(sys.argv[4]) if len(sys.argv) > 4 else 1     # 1: every frame handed over once (sba_descriptors_create), 0: raw tensors per match
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
g = torch.Generator(device=dev); g.manual_seed(7)
base = torch.randn(N, 64, device=dev, generator=g)
frames = []
for f in range(F):     # every frame: the same scene descriptors, shuffled and perturbed, so that matches exist
    perm = torch.randperm(N, device=dev, generator=g)
    d = base[perm] + 0.03 * torch.randn(N, 64, device=dev, generator=g)
    frames.append((d / d.norm(dim=1, keepdim=True)).contiguous())
pairs = sharding.shard_pairs(F, rank, world)
streams = [torch.cuda.Stream(dev) for _ in range(K)]
ctxs = [Context(local, stream=s.cuda_stream) for s in streams]
if PREP:
    frames = [ctxs[f % K].prepare_descriptors(d) for f, d in enumerate(frames)]
for c in ctxs:   # warm-up: scratch buffers of every context
    c.match_begin(frames[0], frames[1]).end()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
pending = [None] * K
n_matches = 0
e0.record(streams[0])
for k, (i, j) in enumerate(pairs):
    s = k % K
    if pending[s] is not None:
        n_matches += len(pending[s].end())
    pending[s] = ctxs[s].match_begin(frames[i], frames[j])
for s in range(K):
    if pending[s] is not None:
        n_matches += len(pending[s].end())
    ends[s].record(streams[s])
torch.cuda.synchronize()
ms = max(e0.elapsed_time(e) for e in ends)
t = torch.tensor([ms, float(len(pairs)), float(n_matches)], dtype=torch.float64, device=dev)
if world > 1:
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM); ms = float(tm[0])
if rank == 0:
    total = int(t[1])
    print(json.dumps({"workload": f"C3: {F} frames x {N} SURF-64 descriptors, all {F * (F - 1) // 2} pairs", "n_gpus": world, "pairs": total,
                      "ms_total": ms, "pairs_per_s": total / (ms * 1e-3), "ms_per_pair_per_gpu": ms * world / total if total else None,
                      "contexts_in_flight": K, "prepared_sets": bool(PREP), "matches_total": int(t[2]), "algorithmic_tflops": 2.0 * 64 * N * N * total / (ms * 1e-3) / 1e12}))
if world > 1:
    dist.destroy_process_group()
