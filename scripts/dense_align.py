#!/usr/bin/env python
"""Profiling harness: the dense registration leg of bench.py on its own
(BASELINE.json configs[2]/[3]), small enough to sit under ncu.

    python scripts/dense_align.py [--src N] [--map N] [--voxel V] [--iters I] [--mode 1|7] [--reps R] [--sort 0|1] [--compact 0|1]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eskf_lio_b200 import capi, synth as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", type=int, default=2_000_000)
    ap.add_argument("--map", type=int, default=10_000_000)
    ap.add_argument("--voxel", type=float, default=0.1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sort", type=int, default=0, help="1: sort the source by voxel brick on the host")
    ap.add_argument("--hint", type=int, default=9_000_000, help="voxel capacity hint of the map")
    ap.add_argument("--compact", type=int, default=0, help="1: eskf_map_compact after the map build")
    a = ap.parse_args()
    ctx = capi.Context(0)
    rng = np.random.default_rng(44)
    scene = S.block_scene()
    gmap = capi.Map(ctx, a.voxel, 1000, a.hint)
    chunk = 2_500_000
    left = a.map
    while left > 0:
        n = min(chunk, left)
        p, c = S.dense_cloud(scene, n, rng)
        gmap.insert(p, c, np.eye(4))
        left -= n
    if a.compact:
        gmap.compact()
    p, c = S.dense_cloud(scene, a.src, rng)
    if a.sort:
        k = np.floor(p / a.voxel).astype(np.int64)
        order = np.lexsort((k[:, 2], k[:, 1], k[:, 0], k[:, 2] >> 2, k[:, 1] >> 2, k[:, 0] >> 2))
        p, c = np.ascontiguousarray(p[order]), np.ascontiguousarray(c[order])
    src = capi.Cloud(ctx, a.src).upload(p, c)
    guess = S.perturbation(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
    for _ in range(a.warmup):
        gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=a.mode)
    times = []
    for _ in range(a.reps):
        ctx.sync()
        ctx.timer_start()
        r = gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=a.mode, trace=True)
        times.append(ctx.timer_stop())
    ms = float(np.median(times))
    per_pt = 136 if a.mode == 1 else 520
    print(json.dumps({"src": a.src, "map_points": a.map, "voxels": gmap.size(), "slots": gmap.capacity(),
                      "voxel": a.voxel, "mode": a.mode, "iters": a.iters, "ms_per_launch": ms,
                      "ms_per_iter": ms / a.iters, "hit_rate": float(r["ncorr"][-1]) / a.src / (7 if a.mode == 7 else 1),
                      "alg_GBps": a.src * per_pt * a.iters / (ms * 1e-3) / 1e9,
                      "Mpts_per_s_per_iter": a.src / (ms / a.iters * 1e-3) / 1e6}))


if __name__ == "__main__":
    main()
