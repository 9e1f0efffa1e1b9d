#!/usr/bin/env python
"""Matcher sweep (BASELINE config 5): N = M = 1k .. 64k SURF-64 descriptors, kNN(k=2) + ratio test.

    python tools/bench_matcher.py                         # 1 GPU
    torchrun --nproc-per-node N tools/bench_matcher.py    # query row-blocks sharded over N GPUs, train set replicated

Per size: device time of the whole match call (prep + distance kernel + re-rank + ratio/compaction) and of the
distance kernel alone, algorithmic TFLOP/s (2*D*N*M) against the measured bf16 peak, for both algorithms."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from spherical_bundle_adjuster_b200 import Context, MATCH_SIMT_EXACT, MATCH_TENSOR, MATCH_TENSOR_FP16, sharding, synth

def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1: dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    sizes = [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    for n in sizes:
        A, B, truth = synth.make_descriptors(n, n, 64, seed=n)
        lo, hi = sharding.shard_range(n, rank, world)             # this rank's query rows
        dA, dB = torch.from_numpy(A[lo:hi]).to(dev), torch.from_numpy(B).to(dev)
        row = {"n": n, "n_gpus": world}
        for name, algo in (("tensor_fp16", MATCH_TENSOR_FP16), ("tensor_bf16x3", MATCH_TENSOR), ("simt", MATCH_SIMT_EXACT)):
            if algo == MATCH_SIMT_EXACT and n > 16384: continue
            ctx.set_profiling(True)
            for _ in range(2): m = ctx.match_two_image(dA, dB, 0.3, algo=algo)
            kms = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            ctx.set_profiling(False)
            if world > 1: dist.barrier()
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps): m = ctx.match_two_image(dA, dB, 0.3, algo=algo)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            ctx.set_profiling(True); ctx.match_two_image(dA, dB, 0.3, algo=algo); kms = ctx.kernel_ms(0); ctx.set_profiling(False)
            t = torch.tensor([ms, kms], device=dev, dtype=torch.float64)
            if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, kms = float(t[0]), float(t[1])
            flops = 2.0 * 64 * n * n
            row[name] = {"call_ms": ms, "kernel_ms": kms, "algorithmic_tflops_call": flops / ms / 1e9, "algorithmic_tflops_kernel": flops / kms / 1e9,
                         "frac_of_bf16_peak_kernel": flops / kms / 1e9 / pk["bf16_tflops"], "matches_this_rank": len(m)}
            planted = np.flatnonzero(truth[lo:hi] >= 0)
            assert len(m) == len(planted)
        if rank == 0: print(json.dumps(row))
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
