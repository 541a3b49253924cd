#!/usr/bin/env python
"""Per-frame stage times of the odometry sequence (debug aid): python scripts/frame_probe.py [n_frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from eskf_lio_b200 import capi, odometry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 45
scans, imu = bench.make_log(n, 43)
od = odometry.Odometry(odometry.default_config(device_resident=1, **bench.odom_overrides()), 0)
ctx = od.context()
clouds = [capi.Cloud(ctx, len(x)).upload_f32(x) for x, _ in scans]
prev = np.zeros(3)
k = 0
for i, (xyz, t) in enumerate(scans):
    end = t[-1]
    while k < imu.shape[0] and imu[k, 0] <= end:
        od.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
        k += 1
        od.spin_once()
    od.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
    k += 1
    od.feed_lidar_cloud(clouds[i], t)
    t0 = time.perf_counter()
    assert od.spin_once()
    dt = 1e3 * (time.perf_counter() - t0)
    inf = od.info()
    cur = np.array(inf.stage_sum_ms)
    print(i, "wall %.3f dev %.3f" % (dt, inf.device_frame_ms_last), "stages", np.round(cur - prev, 3),
          "it", inf.last_iterations, "ins", inf.last_inserted, "vox", inf.map_voxels, flush=True)
    prev = cur
