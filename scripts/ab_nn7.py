#!/usr/bin/env python
"""A/B of the 7-neighbour registration over eskf_ctx options, one map build per voxel size:
python scripts/ab_nn7.py --voxels 0.1,0.5 --cells "align_flags=16;align_flags=4112;align_flags=12304" """
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
    ap.add_argument("--voxels", default="0.1,0.5")
    ap.add_argument("--cells", default="align_flags=16;align_flags=4112;align_flags=12304")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--compact", default="0,1")
    a = ap.parse_args()
    ctx = capi.Context(0)
    for voxel in [float(v) for v in a.voxels.split(",")]:
        rng = np.random.default_rng(44)
        scene = S.block_scene()
        gmap = capi.Map(ctx, voxel, 1000, 9_000_000)
        left = a.map
        while left > 0:
            n = min(2_500_000, left)
            p, c = S.dense_cloud(scene, n, rng)
            gmap.insert(p, c, np.eye(4))
            left -= n
        p, c = S.dense_cloud(scene, a.src, rng)
        src = capi.Cloud(ctx, a.src).upload(p, c)
        guess = S.perturbation(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
        ref = None
        for compact in [int(x) for x in a.compact.split(",")]:
            if compact:
                gmap.compact()
            for cell in a.cells.split(";"):
                for kv in cell.split(","):
                    k, v = kv.split("=")
                    ctx.set_option(k, int(v))
                for _ in range(2):
                    gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=7)
                ts = []
                for _ in range(a.reps):
                    ctx.sync()
                    ctx.timer_start()
                    r = gmap.align_cloud_fixed(src, guess, a.iters, neighbor_mode=7, trace=True)
                    ts.append(ctx.timer_stop())
                if ref is None:
                    ref = r
                print(json.dumps({"voxel": voxel, "compact": compact, "cell": cell,
                                  "ms_per_iter": round(float(np.median(ts)) / a.iters, 4),
                                  "ncorr_equal": bool(np.array_equal(r["ncorr"], ref["ncorr"])),
                                  "max_abs_dT": float(np.abs(r["T"] - ref["T"]).max())}), flush=True)
        del src, gmap
    return 0


if __name__ == "__main__":
    sys.exit(main())
