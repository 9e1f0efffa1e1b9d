#!/usr/bin/env python
"""BA residual+Jacobian evaluation throughput and HBM roofline (BASELINE config 4 and a scaled variant).

    python tools/bench_ba.py                       # 1 GPU: C4 (1M obs, 1024 cams) + 64M-obs variant
    torchrun --nproc-per-node N tools/bench_ba.py  # residual-sharded LM iterations with an NCCL all-reduce

Prints one JSON object per configuration: evals/s, achieved GB/s against the measured HBM peak
(fused path 32 B/observation; materialised residual+Jacobian 80 B/observation), LM iteration time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spherical_bundle_adjuster_b200 import Context, PeerComm, sharding, synth  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def bearings_on_device(n, n_cam, dev, seed):
    """Synthetic bearings generated on the device (64M observations would take minutes in NumPy)."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    b1 = torch.randn((n, 3), generator=g, device=dev, dtype=torch.float32)
    b1 = b1 / b1.norm(dim=1, keepdim=True)
    cam = torch.randint(0, n_cam, (n,), generator=g, device=dev, dtype=torch.int32)
    r_true = 0.3 * torch.randn((n_cam, 3), generator=g, device=dev, dtype=torch.float64)
    th = r_true.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = (r_true / th).float()[cam.long()]
    thc = th.float()[cam.long()]
    b2 = b1 * torch.cos(thc) + torch.cross(k, b1, dim=1) * torch.sin(thc) + k * (k * b1).sum(1, keepdim=True) * (1 - torch.cos(thc))
    b2 = b2 + 1e-3 * torch.randn((n, 3), generator=g, device=dev, dtype=torch.float32)
    b2 = b2 / b2.norm(dim=1, keepdim=True)
    z = torch.zeros((n, 1), device=dev, dtype=torch.float32)
    return torch.cat([b1, z], 1).contiguous(), torch.cat([b2, z], 1).contiguous(), cam, r_true.cpu().numpy()


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    hbm = peaks()
    exchange = os.environ.get("SBA_EXCHANGE", "peer")     # peer (in-kernel NVLink exchange) | nccl (host-launched all-reduce)
    comm = PeerComm(ctx, rank, world, max_cameras=1024) if (world > 1 and exchange == "peer") else None
    configs = [("C4: 1024 cameras, 1M observations (L2-resident: HBM fraction not claimed)", 1_000_000, 1024),
               ("C4 x64: 1024 cameras, 64M observations (2 GB >> L2)", 64_000_000, 1024),
               ("single camera, 16M observations", 16_000_000, 1)]
    for name, n_total, n_cam in configs:
        lo, hi = sharding.shard_range(n_total, rank, world)
        n = hi - lo
        b1, b2, cam, r_true = bearings_on_device(n, n_cam, dev, seed=11 + rank)
        prob = ctx.ba_problem(b1, b2, cam if n_cam > 1 else None, n_cam)
        del b1, b2, cam
        r0 = r_true + 0.02
        out = {"config": name, "n_gpus": world, "obs_per_gpu": n, "exchange": exchange if world > 1 else None}
        if world == 1:
            for mat, bytes_per in ((False, 32), (True, 80)):
                ms = prob.eval_timed(r0, materialise=mat, iters=20)
                gbs = n * bytes_per / (ms * 1e-3) / 1e9
                out["materialised" if mat else "fused"] = {"kernel_ms": ms, "evals_per_s": n / (ms * 1e-3), "bytes_per_obs": bytes_per,
                                                           "achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm}
        elif comm is not None:
            prob.set_comm(comm)
        else:
            prob.set_allreduce(sharding.make_nccl_allreduce(dev))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        prob.solve(r0, max_iter=50)                       # warm-up solve (allocations, NCCL communicator)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r, s = prob.solve(r0, max_iter=50)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["lm"] = {"iterations": s.iterations, "evaluations": s.evaluations, "seconds": float(t[0]), "evals_per_s": n_total * s.evaluations / float(t[0]),
                     "max_err_vs_truth_rad": float(np.abs(r - r_true).max()), "final_cost": s.final_cost}
        if rank == 0:
            print(json.dumps(out))
        prob.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
