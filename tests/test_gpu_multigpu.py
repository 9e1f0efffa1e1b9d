"""GPU (needs >= 2 B200s; skipped otherwise): residual-sharded rotation BA with the per-camera blocks
summed over ranks by NCCL between the evaluation and the decision kernels.  Every rank must return the
same rotations, equal to the single-GPU / oracle solution."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out, exchange="nccl"):
    import torch
    import torch.distributed as dist

    from spherical_bundle_adjuster_b200 import Context, PeerComm, sharding, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n_cam, n = 16, 40000
    b1, b2, cam, r_true = synth.make_bearings(n, noise=1e-3, outlier_frac=0.05, seed=7, n_cam=n_cam)
    b1f, b2f = b1.astype(np.float32), b2.astype(np.float32)
    lo, hi = sharding.shard_range(n, rank, world)
    ctx = Context(rank)
    prob = ctx.ba_problem(b1f[lo:hi], b2f[lo:hi], cam[lo:hi], n_cam)
    comm = None
    if exchange == "peer":      # blocks summed inside the evaluation kernel over NVLink peer memory
        comm = PeerComm(ctx, rank, world, max_cameras=n_cam)
        prob.set_comm(comm)
    else:                       # host-launched NCCL all-reduce between the evaluation and the decision kernels
        prob.set_allreduce(sharding.make_nccl_allreduce(torch.device("cuda", rank)))
    r0 = r_true + 0.05
    r, s = prob.solve(r0)
    ev = prob.eval(r0)                                  # all-reduced blocks of the full problem
    gathered = [None] * world
    dist.all_gather_object(gathered, (r, s.iterations))
    if rank == 0:
        out.put((gathered, ev["H"], ev["cost"], b1f, b2f, cam, r0))
    dist.barrier()
    prob.close()
    if comm is not None:
        comm.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_ba_residual_sharded(exchange):
    import torch
    import torch.multiprocessing as mp

    import oracle
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(k, world, port, q, exchange)) for k in range(world)]
    for p in procs:
        p.start()
    gathered, H, cost, b1f, b2f, cam, r0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r_or, s_or = oracle.ba_rot_solve(b1f.astype(np.float64), b2f.astype(np.float64), cam, r0)
    for r, it in gathered:
        assert np.array_equal(r, gathered[0][0])                 # identical on every rank
        assert np.abs(r - r_or).max() < 1e-6 and it == s_or.iterations
    _, _, H_or, _, c_or = oracle.ba_rot_eval(b1f.astype(np.float64), b2f.astype(np.float64), cam, r0)
    scale = np.abs(H_or).max(axis=1, keepdims=True)
    assert np.all(np.abs(H - H_or) <= 1e-5 * scale) and np.allclose(cost, c_or, rtol=1e-10)


def _worker_module(rank, world, port, out, what):
    """The exact checks bench.py --gpus N runs (spherical_bundle_adjuster_b200/multigpu.py), on small problems."""
    import torch
    import torch.distributed as dist

    from spherical_bundle_adjuster_b200 import Context, multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ctx = Context(rank)
    if what == "ba":
        res = [multigpu.sharded_ba_solve(ctx, rank, world, dev, n_obs, n_cam, reps=2) for n_obs, n_cam in ((200_000, 64), (300_001, 1), (1_000_000, 1024))]
    else:
        res = [multigpu.sharded_match(ctx, rank, world, dev, n, reps=2) for n in (4096, 16384, 20000)]
    if rank == 0:
        out.put(res)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _spawn(what):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker_module, args=(k, world, port, q, what)) for k in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=280)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return world, res


@pytest.mark.timeout(360)
def test_sharded_ba_module_parity():
    """Residual-sharded LM solve (peer-memory exchange and NCCL callback): all ranks bit-equal, <= 1e-6 rad from the
    single-GPU solve of the same data, same iteration count."""
    world, res = _spawn("ba")
    for blk in res:
        assert blk["parity_ok"], blk
        for ex in ("peer", "nccl"):
            assert blk[ex]["ranks_bit_equal"] and blk[ex]["max_abs_diff_vs_single_gpu_rad"] <= 1e-6 and blk[ex]["iterations_equal_single_gpu"], blk


@pytest.mark.timeout(360)
def test_sharded_match_module_parity():
    """Query row-blocks per rank (feature_matcher.cpp:42-59 on each block, train set replicated): the rank-ordered
    concatenation equals the single-GPU list, indices and fp32 distance bits, incl. a size that is not a multiple of the tile."""
    world, res = _spawn("match")
    for blk in res:
        assert blk["parity_ok"] and blk["matches"] > 0, blk
