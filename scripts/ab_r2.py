#!/usr/bin/env python
"""Round-2 A/B of the dense registration kernel (BASELINE.json configs[2]) over run-time options:
one map build per voxel size, every cell a set of eskf_ctx options

    python scripts/ab_r2.py --cells "align_depth=4;align_depth=5;align_depth=5,align_ll=0;..." \
        [--voxels 0.1,0.5] [--compact 0,1] [--src N] [--map N] [--world-sim 1]

Every cell is checked against the static 256-thread kernel of round 1 (per-iteration
correspondence counts identical, pose equal to rounding).  --shard K additionally times the
first 1/K of the source (what one rank of a K-way sharded registration computes per iteration).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eskf_lio_b200 import capi, synth as S  # noqa: E402

DEFAULTS = {"align_block": 0, "align_depth": 0, "align_ticket_chunk": 2, "align_dyn16": 3, "align_dynamic_tiles": 1,
            "align_resident": -1, "align_ll": 1, "align_flags": 16, "align_cons": 0, "align_filter": 1, "l2_persist": 1, "align_fat_points": 1 << 17,
            "align_autotune": 0}


def apply(ctx, cell):
    for k, v in DEFAULTS.items():
        ctx.set_option(k, v)
    for kv in cell.split(","):
        kv = kv.strip()
        if not kv:
            continue
        k, v = kv.split("=")
        ctx.set_option(k, int(v))


def timed(ctx, gmap, src, guess, iters, reps, warm=2):
    for _ in range(warm):
        gmap.align_cloud_fixed(src, guess, iters)
    times = []
    r = None
    for _ in range(reps):
        ctx.sync()
        ctx.timer_start()
        r = gmap.align_cloud_fixed(src, guess, iters, trace=True)
        times.append(ctx.timer_stop())
    return float(np.median(times)) * 1e3 / iters, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", type=int, default=2_000_000)
    ap.add_argument("--map", type=int, default=10_000_000)
    ap.add_argument("--voxels", default="0.1")
    ap.add_argument("--cells", default="align_depth=4;align_depth=5")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--compact", default="0")
    ap.add_argument("--shards", default="", help="comma list of K: also time the first 1/K of the source")
    ap.add_argument("--out", default="")
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    ctx = capi.Context(0)
    rows = []
    cells = [c for c in a.cells.split(";")]
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
        shards = {}
        for k in [int(x) for x in a.shards.split(",") if x]:
            m = a.src // k
            shards[k] = capi.Cloud(ctx, m).upload(p[:m], c[:m])
        guess = S.perturbation(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
        apply(ctx, "align_block=256,align_depth=3,align_dynamic_tiles=0")
        ref = gmap.align_cloud_fixed(src, guess, a.iters, trace=True)
        refs = {k: gmap.align_cloud_fixed(s, guess, a.iters, trace=True) for k, s in shards.items()}
        for compact in [int(x) for x in a.compact.split(",")]:
            if compact:
                gmap.compact()
            for cell in cells:
                ablation = cell.startswith("!")   # (timing-only cells: their results are wrong on purpose)
                cell = cell.lstrip("!")
                apply(ctx, cell)
                us, r = timed(ctx, gmap, src, guess, a.iters, a.reps, a.warm)
                row = {"voxel": voxel, "compact": compact, "slots": gmap.capacity(), "voxels": gmap.size(),
                       "cell": cell, "us_per_iter": round(us, 2),
                       "alg_GBps": round(a.src * 136 / us * 1e-3, 1),
                       "ncorr_equal": ablation or bool(np.array_equal(r["ncorr"], ref["ncorr"])),
                       "max_abs_dT": 0.0 if ablation else float(np.abs(r["T"] - ref["T"]).max())}
                for k, s in shards.items():
                    us_k, rk = timed(ctx, gmap, s, guess, a.iters, a.reps)
                    row[f"shard{k}_us"] = round(us_k, 2)
                    row[f"shard{k}_ok"] = ablation or (bool(np.array_equal(rk["ncorr"], refs[k]["ncorr"])) and
                                                       float(np.abs(rk["T"] - refs[k]["T"]).max()) < 1e-9)
                rows.append(row)
                print(json.dumps(row), flush=True)
        del src, gmap, shards
    bad = [r for r in rows if not r["ncorr_equal"] or r["max_abs_dT"] > 1e-9 or
           any(k.endswith("_ok") and not v for k, v in r.items())]
    print("PARITY", "FAIL" if bad else "OK")
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
