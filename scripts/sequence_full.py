#!/usr/bin/env python
"""BASELINE.json configs[1] in full: an N-frame (default 1000) synthetic odometry
sequence, 10 Hz LiDAR + 400 Hz IMU in the 400 m corridor, through the ROS-free
Odometry / ErrorStateKF host classes with the hot path on the B200, and through
the CPU oracle on the same log.  Reports per-frame latency statistics, map
size / eviction sweeps and the trajectory match GPU-vs-oracle (north_star bar:
1e-5 m / 1e-5 rad) and vs the analytic ground truth.

    python scripts/sequence_full.py [--frames 1000] [--no-oracle] [--out profiles/rX_sequence.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from eskf_lio_b200 import odometry, synth as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    t0 = time.time()
    scans, imu = bench.make_log(a.frames, seed=43)
    t_gen = time.time() - t0
    od = odometry.Odometry(odometry.default_config(device_resident=1, **bench.odom_overrides()), 0)
    dev_ms, removed, voxels, inserted, iters = [], [], [], [], []
    prev = [0.0]

    def feed(i):
        od.feed_lidar(scans[i][0].astype(np.float32), scans[i][1])

    # bench.replay() with a per-frame hook: wrap spin_once through the info counters
    k = 0
    poses = []
    wall = []
    for i, (xyz, t) in enumerate(scans):
        end = t[-1]
        while k < imu.shape[0] and imu[k, 0] <= end:
            od.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
            od.spin_once()
        od.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
        k += 1
        w0 = time.perf_counter()
        feed(i)
        assert od.spin_once()
        wall.append(1e3 * (time.perf_counter() - w0))
        inf = od.info()
        poses.append(od.pose())
        if i > 0:
            dev_ms.append(inf.device_frame_ms_last)
        removed.append(int(inf.last_removed))
        voxels.append(int(inf.map_voxels))
        inserted.append(int(inf.last_inserted))
        iters.append(int(inf.last_iterations))
    info = od.info()
    dev = np.array(dev_ms)
    tr = S.corridor_trajectory()
    G0 = tr.pose_world(scans[0][1][-1])
    gt = [np.linalg.inv(G0) @ tr.pose_world(t[-1]) for _, t in scans]
    gt_err = [bench.pose_delta(g, p) for g, p in zip(gt, poses)]
    out = {"frames": a.frames, "generation_s": t_gen,
           "device_ms_per_frame": {"mean": float(dev.mean()), "p50": float(np.percentile(dev, 50)),
                                   "p99": float(np.percentile(dev, 99)), "max": float(dev.max())},
           "e2e_wall_ms_per_frame": {"mean": float(np.mean(wall[1:])), "p50": float(np.percentile(wall[1:], 50)),
                                     "p99": float(np.percentile(wall[1:], 99)), "max": float(np.max(wall[1:]))},
           "stage_avg_ms": list(info.stage_avg_ms), "stage_max_ms": list(info.stage_max_ms),
           "map_voxels_final": voxels[-1], "map_voxels_max": max(voxels),
           "eviction_sweeps": int(np.count_nonzero(np.diff([0] + removed))),
           "voxels_evicted_last_sweep": removed[-1], "frames_inserted": int(sum(inserted)),
           "filter_states": int(info.n_states),
           "vs_ground_truth": {"max_m": max(e[0] for e in gt_err), "max_rad": max(e[1] for e in gt_err),
                               "final_m": gt_err[-1][0]},
           "distance_travelled_m": float(np.linalg.norm(poses[-1][:3, 3])),
           "frames_over_1ms_device": int((dev > 1.0).sum()),
           # the slowest frames: (frame, device ms, e2e wall ms, GN iterations, eviction sweep in this frame?)
           "slowest_frames": [{"frame": int(i + 1), "device_ms": float(dev[i]), "wall_ms": float(wall[i + 1]),
                               "gn_iterations": iters[i + 1],
                               "eviction_sweep": bool(removed[i + 1] != removed[i])}
                              for i in np.argsort(-dev)[:6]],
           "gn_iterations": {"mean": float(np.mean(iters[1:])), "max": int(np.max(iters[1:]))}}
    od.close()
    if not a.no_oracle:
        import oracle as O
        O.build()
        O.set_num_threads(len(os.sched_getaffinity(0)))
        oo = O.Odometry(O.odom_default_config(**bench.odom_overrides()))
        t1 = time.time()
        _, oposes, _ = bench.replay(oo, scans, imu, lambda i: oo.feed_lidar(scans[i][0], scans[i][1]), 1)
        oi = oo.info()
        d = [bench.pose_delta(x, y) for x, y in zip(oposes, poses)]
        dm = np.array([x[0] for x in d])
        out["cpu_oracle"] = {"seconds": time.time() - t1, "cores": O.num_threads(),
                             "stage_avg_ms": list(oi.stage_avg_ms), "ms_per_frame": float(sum(oi.stage_avg_ms)),
                             "map_voxels_final": int(oi.map_voxels)}
        out["trajectory_match_vs_cpu"] = {"max_m": float(dm.max()), "rms_m": float(np.sqrt((dm ** 2).mean())),
                                          "max_rad": max(x[1] for x in d),
                                          "frames_within_1e-5": int(np.count_nonzero(dm < 1e-5))}
    print(json.dumps(out), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
