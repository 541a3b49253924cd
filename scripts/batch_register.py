#!/usr/bin/env python
"""BASELINE.json configs[4]: batched multi-session throughput — independent
scan / local-map registration pairs, one shard of the batch per GPU (no
collective), several pairs in flight per GPU on separate contexts (= CUDA
streams; one 15k-point registration occupies ~60 of the 148 SMs).

    python [-m torch.distributed.run --nproc-per-node N ...] scripts/batch_register.py \
        [--pairs 512] [--streams 4] [--map-points 300000] [--scan-points 15000]

Each pair: a map of --map-points surface samples of the hall scene (0.5 m
voxels) seen from its own station, a --scan-points source cloud and the
config-1 guess perturbation; default ICP parameters, run to convergence.
Reports registrations / s (max over ranks of the wall time of the timed phase;
maps are built beforehand) and, on rank 0, the CPU oracle's rate on a sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eskf_lio_b200 import capi, sharded, synth as S  # noqa: E402


def make_pair(k, map_points, scan_points):
    rng = np.random.default_rng(1000 + k)
    scene = S.hall_scene()
    mp, mc = S.dense_cloud(scene, map_points, rng)
    sp, sc = S.dense_cloud(scene, scan_points, rng)
    # station k: the scan is expressed in a body frame at a pose on the config-1 arc
    T = S.arc_trajectory(1, start=(-8.0 + 0.03 * (k % 400), -3.0 + 0.01 * (k % 37)))[0]
    Ti = np.linalg.inv(T)
    sp_b = sp @ Ti[:3, :3].T + Ti[:3, 3]
    sc_b = Ti[:3, :3] @ sc @ Ti[:3, :3].T
    return mp, mc, np.ascontiguousarray(sp_b), np.ascontiguousarray(sc_b), T @ S.perturbation()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=512)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--map-points", type=int, default=300_000)
    ap.add_argument("--scan-points", type=int, default=15_000)
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--pipelined", type=int, default=1,
                    help="1: one host thread keeps one registration in flight per context (begin/end); "
                         "0: one blocking host thread per context; "
                         "2: the same pipelining inside the library, one eskf_align_batch call per pass")
    ap.add_argument("--repeat", type=int, default=8, help="timed passes over the batch (pairs are independent)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = list(sharded.shard_batch(a.pairs, rank, world))
    ctxs = [capi.Context(local) for _ in range(a.streams)]
    jobs = [[] for _ in range(a.streams)]
    cpu = []
    for n, k in enumerate(mine):
        mp, mc, sp, sc, guess = make_pair(k, a.map_points, a.scan_points)
        ctx = ctxs[n % a.streams]
        gmap = capi.Map(ctx, 0.5, 1000, 1 << 15)
        gmap.insert(mp, mc, np.eye(4))
        cloud = capi.Cloud(ctx, len(sp)).upload(sp, sc)
        jobs[n % a.streams].append((gmap, cloud, guess))
        if rank == 0 and len(cpu) < a.cpu_sample:
            cpu.append((mp, mc, sp, sc, guess))
    for c in ctxs:
        c.sync()
    results = [[] for _ in range(a.streams)]

    def work(s):
        for rep in range(a.repeat):
            for gmap, cloud, guess in jobs[s]:
                r = gmap.align_cloud(cloud, guess)
                if rep == 0:
                    results[s].append(r)

    def work_pipelined():
        """One host thread, one registration in flight per context (eskf_align_cloud_begin / _end):
        the launches of the other contexts overlap the kernel each end() waits for."""
        for rep in range(a.repeat):
            pending = [None] * a.streams
            for k in range(max(len(j) for j in jobs)):
                for s in range(a.streams):
                    if pending[s] is not None:
                        r = ctxs[s].align_end()
                        if rep == 0:
                            results[s].append(r)
                        pending[s] = None
                    if k < len(jobs[s]):
                        gmap, cloud, guess = jobs[s][k]
                        gmap.align_cloud_begin(cloud, guess)
                        pending[s] = True
            for s in range(a.streams):
                if pending[s] is not None:
                    r = ctxs[s].align_end()
                    if rep == 0:
                        results[s].append(r)

    for s in range(a.streams):       # warm-up: first job of every stream
        if jobs[s]:
            jobs[s][0][0].align_cloud(jobs[s][0][1], jobs[s][0][2])
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    if a.pipelined == 2:
        flat = [jobs[n % a.streams][n // a.streams] for n in range(len(mine))]
        for rep in range(a.repeat):
            rs = capi.align_batch(ctxs, [j[0] for j in flat], [j[1] for j in flat], [j[2] for j in flat])
            if rep == 0:
                for n, r in enumerate(rs):
                    results[n % a.streams].append(r)
    elif a.pipelined:
        work_pipelined()
    else:
        threads = [threading.Thread(target=work, args=(s,)) for s in range(a.streams)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    for c in ctxs:
        c.sync()
    dt = time.perf_counter() - t0
    if dist is not None:
        import torch
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
    if rank == 0:
        its = [r["iterations"] for rs in results for r in rs]
        out = {"pairs": a.pairs, "world": world, "streams_per_gpu": a.streams, "pipelined": a.pipelined, "map_points": a.map_points,
               "scan_points": a.scan_points, "passes": a.repeat, "seconds": dt,
               "registrations_per_s": a.pairs * a.repeat / dt,
               "gn_iterations_mean": float(np.mean(its)), "all_converged": all(r["converged"] for rs in results for r in rs)}
        if cpu:
            import oracle as O
            O.build()
            O.set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1
            t_cpu = 0.0
            worst = 0.0
            for i, (mp, mc, sp, sc, guess) in enumerate(cpu):
                om = O.Map(0.5, 1000)
                om.insert(mp, mc)
                t1 = time.perf_counter()
                ro = om.align(sp, sc, guess)
                t_cpu += time.perf_counter() - t1
                rg = results[i % a.streams][i // a.streams]
                E = np.linalg.inv(ro["T"]) @ rg["T"]
                worst = max(worst, float(np.linalg.norm(E[:3, 3])))
                assert ro["iterations"] == rg["iterations"], (i, ro["iterations"], rg["iterations"])
            out["cpu_registrations_per_s"] = len(cpu) / t_cpu
            out["cpu_cores"] = O.num_threads()
            out["cpu_sample"] = len(cpu)
            out["max_pose_delta_vs_cpu_m"] = worst
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
