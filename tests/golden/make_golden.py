#!/usr/bin/env python
"""Generates tests/golden/hotpath_v1.npz from the CPU oracle.

The reference ships no golden vectors and cannot be built here (parity
unpinned), so these fixtures freeze the ORACLE's outputs on a small seeded
case: they guard the oracle against regressions on the CPU and give the GPU
tests a checker that does not depend on liboracle.so being rebuilt.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from eskf_lio_b200 import synth as S  # noqa: E402


def main():
    O.build()
    O.set_num_threads(1)  # thread-count independent summation order for the fixture
    rng = np.random.default_rng(2024)
    scene = S.hall_scene()
    poses = S.arc_trajectory(3)
    T_il = S.default_T_il()
    out = {"T_il": T_il, "poses": np.stack(poses), "voxel": np.float64(0.5)}
    om = O.Map(0.5, 1000)
    for k, T in enumerate(poses):
        xyz, t = S.make_scan(scene, T, rng)
        xyz, t = xyz[::16].copy(), t[::16].copy()
        out[f"raw{k}"] = xyz.astype(np.float32)
        out[f"time{k}"] = t
        p, c, src = O.preprocess(xyz, t, T_il, None, 0.5)
        out[f"kept{k}"] = src
        out[f"cov{k}"] = c
        if k < 2:
            om.update(p, c, T, initialize=True)
        else:
            guess = T @ S.perturbation()
            out["guess"] = guess
            pg, cg = O.transform_cloud(p, c, guess)
            out["keys_at_guess"] = O.voxel_index(pg, 0.5)
            H, b, hit, nc = om.linearize(pg, cg)
            out["lin_H"], out["lin_b"], out["lin_hit"] = H, b, hit[:, 0]
            r = om.align(p, c, guess)
            out["align_T"] = r["T"]
            out["align_iterations"] = np.int64(r["iterations"])
            out["align_ncorr"] = r["ncorr"]
            out["align_H"], out["align_b"] = r["H"], r["b"]
    keys, count, mean, cov = om.export()
    out["map_keys"], out["map_count"], out["map_mean"], out["map_cov"] = keys, count, mean, cov
    path = os.path.join(ROOT, "tests", "golden", "hotpath_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
