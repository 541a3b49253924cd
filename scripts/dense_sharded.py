#!/usr/bin/env python
"""BASELINE.json configs[2]: ONE dense registration sharded over the GPUs of a node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dense_sharded.py [--src 2000000] [--map 10000000] [--voxel 0.1]

One process per GPU.  The voxel map is replicated (every rank inserts the same
seeded points), the source cloud is split by contiguous point range, and every
Gauss-Newton iteration the 28 H/b sums of all ranks are combined
  p2p   inside the persistent kernel, through peer-mapped NVLink mailboxes
        (eskf_align_cloud_p2p: no collective call, no host round trip), or
  nccl  by an NCCL all-reduce between two kernel launches per iteration
        (eskf_align_cloud_sharded + torch.distributed: the baseline).
Times are CUDA events on each rank's stream, max over ranks.  Rank 0 prints one
JSON line; it also checks that every rank returned the bit-identical pose and
that the pose equals the unsharded single-GPU result within 1e-5 m / 1e-5 rad.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eskf_lio_b200 import capi, sharded, synth as S  # noqa: E402


def pose_delta(A, B):
    E = np.linalg.inv(A) @ B
    R = E[:3, :3]
    sin = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.linalg.norm(E[:3, 3])), float(np.arctan2(sin, 0.5 * (np.trace(R) - 1.0)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", type=int, default=2_000_000)
    ap.add_argument("--map", type=int, default=10_000_000)
    ap.add_argument("--voxel", type=float, default=0.1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-device", action="store_true", help="all ranks on cuda:0 (functional test)")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: --src points PER RANK (each rank samples its own range; no unsharded check)")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = 0 if a.same_device else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if a.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(a.backend)

    ctx = capi.Context(local)
    rng = np.random.default_rng(44)
    scene = S.block_scene()
    gmap = capi.Map(ctx, a.voxel, 1000, max(1 << 16, int(0.9 * a.map)))
    left = a.map
    while left > 0:
        n = min(2_500_000, left)
        p, c = S.dense_cloud(scene, n, rng)
        gmap.insert(p, c, np.eye(4))
        left -= n
    if a.weak:
        p, c = S.dense_cloud(scene, a.src, np.random.default_rng(4400 + rank))
        b, e = 0, a.src
    else:
        p, c = S.dense_cloud(scene, a.src, rng)
        b, e = sharded.shard_range(a.src, rank, world)
    total_src = a.src * world if a.weak else a.src
    shard = capi.Cloud(ctx, max(e - b, 64)).upload(p[b:e], c[b:e])
    guess = S.perturbation(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
    comm = sharded.make_comm(ctx) if world > 1 else capi.Comm(ctx, 0, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(2):
            barrier()
            fn()
        ts = []
        r = None
        for _ in range(a.reps):
            barrier()
            ctx.timer_start()
            r = fn()
            ts.append(ctx.timer_stop())
        ms = float(np.median(ts))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, r

    out = {"world": world, "src": total_src, "scaling": "weak" if a.weak else "strong",
           "map_points": a.map, "voxel": a.voxel, "iters": a.iters,
           "voxels": gmap.size(), "shard_points": e - b}
    ms_p2p, r_p2p = timed(lambda: gmap.align_cloud_p2p(shard, guess, comm, fixed_iterations=a.iters, trace=True))
    out["p2p_ms_per_iter"] = ms_p2p / a.iters
    out["p2p_mpts_per_s_per_iter"] = total_src / (ms_p2p / a.iters * 1e-3) / 1e6
    if world > 1 and a.backend == "nccl":
        cb = sharded.TorchAllReduce()
        ms_nccl, r_nccl = timed(lambda: gmap.align_cloud_sharded(shard, guess, cb, fixed_iterations=a.iters))
        out["nccl_ms_per_iter"] = ms_nccl / a.iters
        out["nccl_vs_p2p_pose_delta"] = pose_delta(r_nccl["T"], r_p2p["T"])
    # run to convergence through the fused path (parity leg)
    barrier()
    r_conv = gmap.align_cloud_p2p(shard, guess, comm, trace=True)
    out["converged_iterations"] = r_conv["iterations"]
    # every rank must hold the bit-identical pose
    same = True
    if world > 1:
        t = torch.tensor(r_conv["T"].ravel(), dtype=torch.float64, device=f"cuda:{local}")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    out["identical_pose_on_all_ranks"] = same
    if rank == 0 and not a.weak:
        full = capi.Cloud(ctx, a.src).upload(p, c)
        ref = gmap.align_cloud(full, guess, trace=True)
        out["unsharded_iterations"] = ref["iterations"]
        out["vs_unsharded_pose_delta"] = pose_delta(ref["T"], r_conv["T"])
        nit = min(len(ref["H"]), len(r_conv["H"]))
        out["vs_unsharded_H_rel"] = float(max(
            np.linalg.norm(r_conv["H"][k] - ref["H"][k]) / np.linalg.norm(ref["H"][k]) for k in range(nit)))
        out["ncorr_equal"] = bool(np.array_equal(ref["ncorr"][:nit], r_conv["ncorr"][:nit]))
    barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
