#!/usr/bin/env python
"""BASELINE.json configs[3]: voxel-resolution x neighbourhood sweep of the
registration kernel on the dense config-3 data (2M-point source vs a 10M-point
map): per cell ms / GN iteration, Mpts/s, hit rate, algorithmic GB/s
(136 B/pt/iter for 1 neighbour, 520 B/pt/iter upper bound for 7)
and the hit-weighted figure.  One JSON line per cell + a markdown table.

    python scripts/sweep.py [--src N] [--map N] [--out profiles/rX_sweep.md]
For the DRAM / L2 counters of a cell run scripts/dense_align.py under ncu.
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--voxels", default="0.1,0.25,0.5,1.0")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ctx = capi.Context(0)
    scene = S.block_scene()
    rng = np.random.default_rng(44)
    chunks = []
    left = a.map
    while left > 0:
        n = min(2_500_000, left)
        chunks.append(S.dense_cloud(scene, n, rng))
        left -= n
    p, c = S.dense_cloud(scene, a.src, rng)
    src = capi.Cloud(ctx, a.src).upload(p, c)
    rows = []
    for voxel in [float(v) for v in a.voxels.split(",")]:
        gmap = capi.Map(ctx, voxel, 1000, 1 << 20)
        for mp, mc in chunks:
            gmap.insert(mp, mc, np.eye(4))
        # perturbation inside the voxel basin (SURVEY.md 8d config 3 scales it with the voxel)
        guess = S.perturbation(dt=(0.3 * voxel, -0.15 * voxel, 0.1 * voxel), angle_deg=0.3)
        # two table layouts: as the bulk inserts grew it (sized for "every incoming point a new
        # voxel": load factor 0.03-0.25) and after eskf_map_compact (load factor 1/2).  Records
        # live in the table slots, so the layout sets the address range the gathers spread over.
        for layout in ("grown", "compact"):
            if layout == "compact":
                gmap.compact()
            for mode in (1, 7):
                for _ in range(2):
                    gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=mode)
                ts = []
                for _ in range(a.reps):
                    ctx.sync()
                    ctx.timer_start()
                    r = gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=mode, trace=True)
                    ts.append(ctx.timer_stop())
                ms = float(np.median(ts)) / a.iters
                hits = float(np.mean(r["ncorr"]))
                per_pt = 136 if mode == 1 else 520
                row = {"voxel": voxel, "neighbors": mode, "map_voxels": gmap.size(), "layout": layout,
                       "slots": gmap.capacity(), "ms_per_iter": ms,
                       "mpts_per_s": a.src / (ms * 1e-3) / 1e6, "hits_per_point": hits / a.src,
                       "alg_GBps": a.src * per_pt / (ms * 1e-3) / 1e9,
                       "hit_weighted_GBps": (48.0 * a.src + 24.0 * min(hits, a.src) + 64.0 * hits) / (ms * 1e-3) / 1e9}
                rows.append(row)
                print(json.dumps(row), flush=True)
        gmap.close()
    if a.out:
        with open(a.out, "w") as f:
            f.write(f"# voxel-size x neighbourhood sweep (configs[3]): {a.src} source pts vs {a.map}-pt map, "
                    f"{a.iters} GN iterations per launch, median of {a.reps}\n\n"
                    "| voxel m | neighbours | map voxels | table | slots | ms / GN iter | Mpts/s | hits / point | "
                    "algorithmic GB/s | hit-weighted GB/s |\n|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['voxel']} | {r['neighbors']} | {r['map_voxels']} | {r['layout']} | {r['slots']} | "
                        f"{r['ms_per_iter']:.4f} | "
                        f"{r['mpts_per_s']:.0f} | {r['hits_per_point']:.3f} | {r['alg_GBps']:.0f} | "
                        f"{r['hit_weighted_GBps']:.0f} |\n")


if __name__ == "__main__":
    main()
