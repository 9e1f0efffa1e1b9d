"""The two sharded paths of SURVEY section 8(e) run under ``torch.distributed`` (one process per GPU), each
with its parity assertion built in.  bench.py calls them after the headline timing when WORLD_SIZE > 1 and
``tests/test_gpu_multigpu.py`` spawns them on boxes with >= 2 GPUs.

* :func:`sharded_ba_solve` -- BASELINE config 4: the observations of ONE bundle-adjustment problem are cut into
  contiguous shards (any camera may appear in any shard); every LM iteration the per-camera normal-equation
  blocks ``[n_cam x 10]`` are summed over ranks, either inside the evaluation kernel over NVLink peer memory
  (``PeerComm``) or by a host-launched NCCL all-reduce between the evaluation and the decision kernels.
  Replaces the single ceres::Solve of spherical_bundle_adjuster.cpp:202-203 over the residuals added by
  ba_spherical_costfunctor_rot_only::add_residual (:921-945).
* :func:`sharded_match` -- BASELINE config 5: query row-blocks of one big feature_matcher::match_two_image
  (feature_matcher.cpp:42-59) go to different GPUs, the train set is replicated, the per-rank match lists
  concatenate in rank order (= ascending queryIdx).  No collective on the data path; the gather of the lists
  is the host-side result hand-over.

All timings are device times (CUDA events on the stream the library context runs on), maximum over ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .api import Context, PeerComm


def synth_bearings_device(n: int, n_cam: int, dev, seed: int):
    """Seeded bearings generated on the device (64 M observations take minutes in NumPy).  The SAME seed gives
    the SAME tensors on every rank, so each rank can slice its shard out of the full problem."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    b1 = torch.randn((n, 3), generator=g, device=dev, dtype=torch.float32)
    b1 = b1 / b1.norm(dim=1, keepdim=True)
    cam = torch.randint(0, n_cam, (n,), generator=g, device=dev, dtype=torch.int32)
    r_true = 0.3 * torch.randn((n_cam, 3), generator=g, device=dev, dtype=torch.float64)
    th = r_true.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = (r_true / th).float()[cam.long()]
    thc = th.float()[cam.long()]
    b2 = b1 * torch.cos(thc) + torch.cross(k, b1, dim=1) * torch.sin(thc) + k * (k * b1).sum(1, keepdim=True) * (1 - torch.cos(thc))
    del k, thc
    b2 = b2 + 1e-3 * torch.randn((n, 3), generator=g, device=dev, dtype=torch.float32)
    b2 = b2 / b2.norm(dim=1, keepdim=True)
    z = torch.zeros((n, 1), device=dev, dtype=torch.float32)
    return torch.cat([b1, z], 1).contiguous(), torch.cat([b2, z], 1).contiguous(), cam, r_true.cpu().numpy()


def _max_over_ranks(x: float, dev) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _timed_solve(prob, r0, dev, reps: int):
    """(r, summary, device ms per solve): the solve is timed `reps` times after one warm-up solve."""
    prob.solve(r0, max_iter=50)                                   # warm-up: allocations, NCCL communicator, peer mappings
    torch.cuda.synchronize(dev)
    if dist.is_initialized():
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r, s = prob.solve(r0, max_iter=50)
    e1.record()
    torch.cuda.synchronize(dev)
    return r, s, e0.elapsed_time(e1) / reps


def sharded_ba_solve(ctx: Context, rank: int, world: int, dev, n_obs: int, n_cam: int, seed: int = 11, reps: int = 3,
                     exchanges=("peer", "nccl")) -> dict:
    """One rotation-only LM solve of an (n_obs, n_cam) problem with its residuals sharded over `world` ranks, once
    per exchange mechanism.  Asserts (on every rank) that all ranks end with bit-identical rotations, and (rank 0)
    that they agree with the single-GPU solve of the SAME data to 1e-6 rad with the same iteration count.
    `ctx` must run on torch's current stream (the NCCL callback enqueues the collective there)."""
    b1, b2, cam, r_true = synth_bearings_device(n_obs, n_cam, dev, seed)
    r0 = r_true + 0.02
    lo, hi = sharding.shard_range(n_obs, rank, world)
    out = {"n_obs": n_obs, "n_cam": n_cam, "obs_per_rank": hi - lo, "world": world, "parity_ok": True}

    # the N=1 answer: rank 0 solves the whole problem alone
    ref = None
    if rank == 0:
        p1 = ctx.ba_problem(b1, b2, cam if n_cam > 1 else None, n_cam)
        r1, s1, ms1 = _timed_solve_local(p1, r0, dev, reps)
        p1.close()
        ref = (r1, s1.iterations, s1.evaluations, s1.final_cost)
        out["single_gpu"] = {"ms_per_solve": ms1, "us_per_evaluation": 1e3 * ms1 / max(1, s1.evaluations), "iterations": s1.iterations,
                             "evaluations": s1.evaluations, "max_err_vs_truth_rad": float(np.abs(r1 - r_true).max())}
    if world > 1:
        dist.barrier()

    sb1, sb2, scam = b1[lo:hi].contiguous(), b2[lo:hi].contiguous(), cam[lo:hi].contiguous()
    del b1, b2, cam
    prob = ctx.ba_problem(sb1, sb2, scam if n_cam > 1 else None, n_cam)
    comm = None
    for ex in exchanges:
        if world == 1:
            break
        if ex == "peer":
            if comm is None:
                comm = PeerComm(ctx, rank, world, max_cameras=n_cam)
            prob.set_comm(comm)
        else:
            prob.set_comm(None)
            prob.set_allreduce(sharding.make_nccl_allreduce(dev))
        r, s, ms = _timed_solve(prob, r0, dev, reps)
        ms = _max_over_ranks(ms, dev)
        # every rank must hold the same bits
        mine = torch.from_numpy(np.ascontiguousarray(r)).to(dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        same = all(torch.equal(a, allr[0]) for a in allr)
        blk = {"ms_per_solve": ms, "us_per_evaluation": 1e3 * ms / max(1, s.evaluations), "iterations": s.iterations,
               "evaluations": s.evaluations, "ranks_bit_equal": bool(same)}
        ok = same
        if rank == 0:
            err = float(np.abs(r - ref[0]).max())
            blk["max_abs_diff_vs_single_gpu_rad"] = err
            blk["iterations_equal_single_gpu"] = bool(s.iterations == ref[1])
            blk["rel_cost_diff_vs_single_gpu"] = float(abs(s.final_cost - ref[3]) / max(abs(ref[3]), 1e-300))
            ok = ok and err <= 1e-6 and s.iterations == ref[1]
        out[ex] = blk
        out["parity_ok"] = bool(out["parity_ok"] and ok)
        if ex == "nccl":
            prob.set_allreduce(None)
    prob.close()
    if comm is not None:
        comm.close()
    flag = torch.tensor([1 if out["parity_ok"] else 0], device=dev, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["parity_ok"] = bool(int(flag[0]) == 1)
    return out


def _timed_solve_local(prob, r0, dev, reps: int):
    """Like _timed_solve but without barriers (only rank 0 runs it)."""
    prob.solve(r0, max_iter=50)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r, s = prob.solve(r0, max_iter=50)
    e1.record()
    torch.cuda.synchronize(dev)
    return r, s, e0.elapsed_time(e1) / reps


def sharded_match(ctx: Context, rank: int, world: int, dev, n: int, dim: int = 64, seed: int | None = None, ratio: float = 0.3,
                  reps: int = 5) -> dict:
    """match_two_image of n x n descriptors with the query rows cut into `world` contiguous row-blocks.  The per-rank
    lists are gathered in rank order and rank 0 asserts they equal the list of ONE single-GPU call on the whole
    query set: same (queryIdx, trainIdx) and the same fp32 distance bits."""
    from . import synth
    A, B, _ = synth.make_descriptors(n, n, dim, seed=n if seed is None else seed)
    lo, hi = sharding.shard_range(n, rank, world)
    dA, dB = torch.from_numpy(A[lo:hi]).to(dev), torch.from_numpy(B).to(dev)
    for _ in range(2):
        m = ctx.match_two_image(dA, dB, ratio)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m = ctx.match_two_image(dA, dB, ratio)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = _max_over_ranks(e0.elapsed_time(e1) / reps, dev)
    mine = (m.query_idx.cpu().numpy().astype(np.int64) + lo, m.train_idx.cpu().numpy().astype(np.int64), m.distance.cpu().numpy())
    lists = [None] * world
    if world > 1:
        dist.all_gather_object(lists, mine)
    else:
        lists = [mine]
    out = {"n": n, "dim": dim, "world": world, "rows_per_rank": hi - lo, "ms_per_call": ms,
           "algorithmic_tflops": 2.0 * dim * n * n / (ms * 1e-3) / 1e12, "parity_ok": True}
    if rank == 0:
        qi = np.concatenate([x[0] for x in lists]); ti = np.concatenate([x[1] for x in lists]); dd = np.concatenate([x[2] for x in lists])
        fullA = torch.from_numpy(A).to(dev)
        for _ in range(2):
            full = ctx.match_two_image(fullA, dB, ratio)
        torch.cuda.synchronize(dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(reps):
            full = ctx.match_two_image(fullA, dB, ratio)
        f1.record()
        torch.cuda.synchronize(dev)
        fq, ft, fd = full.query_idx.cpu().numpy().astype(np.int64), full.train_idx.cpu().numpy().astype(np.int64), full.distance.cpu().numpy()
        ok = (len(qi) == len(fq) and np.array_equal(qi, fq) and np.array_equal(ti, ft)
              and np.array_equal(dd.view(np.uint32), fd.view(np.uint32)) and bool(np.all(np.diff(qi) > 0)))
        out.update(matches=int(len(qi)), single_gpu_ms_per_call=f0.elapsed_time(f1) / reps, parity_ok=bool(ok))
    flag = torch.tensor([1 if out["parity_ok"] else 0], device=dev, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["parity_ok"] = bool(int(flag[0]) == 1)
    return out
