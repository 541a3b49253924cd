#!/usr/bin/env python
"""Generates tests/golden/odometry_v1.npz from the CPU oracle (the reference ships
no vectors and cannot be built here): a short synthetic 10 Hz LiDAR + 400 Hz IMU log
(decimated sweeps so the file stays small) and the oracle's per-frame poses, Gauss-
Newton iteration counts, keyframe-gate decisions and final filter state.
    python tests/golden/make_golden_odometry.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from eskf_lio_b200 import synth as S  # noqa: E402

N_FRAMES = 20
DECIM = 16


def main():
    O.build()
    O.set_num_threads(1)  # thread-count independent summation order for the fixture
    tr = S.hall_trajectory()
    scans, imu = S.make_sequence(S.hall_scene(), tr, N_FRAMES, seed=77)
    scans = [(x[::DECIM].copy(), t[::DECIM].copy()) for x, t in scans]
    od = O.Odometry(O.odom_default_config(map_voxel_size=0.5, preprocess_voxel_size=0.5))
    rec = []
    poses = O.run_sequence(od, scans, imu, lambda i, o: rec.append(
        (o.info().last_iterations, o.info().last_inserted, o.info().map_voxels, o.info().last_kept)))
    st = od.last_state(with_P=True)
    out = {"imu": imu, "n": np.array([len(x) for x, _ in scans]),
           "xyz": np.concatenate([x for x, _ in scans]).astype(np.float32),
           "time": np.concatenate([t for _, t in scans]),
           "poses": np.stack(poses), "rec": np.array(rec, dtype=np.int64),
           "state": np.concatenate([[st["t"]], st["p"], st["v"], st["q"], st["ba"], st["bg"], st["g"]]),
           "P": st["P"]}
    path = os.path.join(ROOT, "tests", "golden", "odometry_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
