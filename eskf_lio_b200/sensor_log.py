"""Flat binary sensor log ("ESKFLOG1"): the wire format of the reference's two ROS subscribers
without ROS (include/ESKF_LIO/Subscriber.hpp:38-52 sensor_msgs/Imu; :80-103 sensor_msgs/PointCloud2
with float32 x, y, z and a float64 "timestamp" per point), records in callback order.  The C++
reader / writer is eskf_lio_b200/host/ESKF_LIO/SensorLog.hpp (eskf_odom_replay_log replays a file
through the odometry); this module is its NumPy mirror for tools and tests.

    header = b"ESKFLOG1"  u32 version (1)  u32 0
    record = u32 type  u32 count  payload        (little endian)
      type 1, count 1:  f64 stamp, f64 angular_velocity[3], f64 linear_acceleration[3]
      type 2:           count x {f32 x, f32 y, f32 z, f64 timestamp}   (20 B, packed)
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"ESKFLOG1"
IMU, LIDAR = 1, 2
POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<f8")])  # packed: 20 bytes
assert POINT.itemsize == 20


class Writer:
    def __init__(self, path: str):
        self.f = open(path, "wb")
        self.f.write(MAGIC + struct.pack("<II", 1, 0))

    def imu(self, stamp: float, gyro, acc):
        self.f.write(struct.pack("<II7d", IMU, 1, float(stamp), *map(float, gyro), *map(float, acc)))

    def lidar(self, xyz, point_time):
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
        rec = np.empty(len(xyz), dtype=POINT)
        rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        rec["t"] = np.asarray(point_time, dtype=np.float64)
        self.f.write(struct.pack("<II", LIDAR, len(rec)))
        self.f.write(rec.tobytes())

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read(path: str):
    """Yields ("imu", stamp, gyro[3], acc[3]) and ("lidar", xyz float32 [n, 3], point_time [n])."""
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) != 16 or head[:8] != MAGIC:
            raise ValueError(f"{path} is not an ESKFLOG1 file")
        if struct.unpack("<II", head[8:])[0] != 1:
            raise ValueError("unsupported log version")
        while True:
            h = f.read(8)
            if not h:
                return
            if len(h) != 8:
                raise ValueError("truncated record header")
            typ, count = struct.unpack("<II", h)
            if typ == IMU:
                b = f.read(56)
                if len(b) != 56 or count != 1:
                    raise ValueError("bad IMU record")
                v = struct.unpack("<7d", b)
                yield "imu", v[0], np.array(v[1:4]), np.array(v[4:7])
            elif typ == LIDAR:
                b = f.read(20 * count)
                if len(b) != 20 * count:
                    raise ValueError("truncated sweep")
                rec = np.frombuffer(b, dtype=POINT)
                yield "lidar", np.stack([rec["x"], rec["y"], rec["z"]], axis=1), rec["t"].copy()
            else:
                raise ValueError(f"unknown record type {typ}")


def write_sequence(path: str, scans, imu):
    """A synthetic log (synth.make_sequence) in callback order: every IMU sample up to and including the
    first one past a sweep's last point, then the sweep (the order bench.py / odometry.run_sequence
    deliver them in), the remaining IMU samples at the end."""
    k = 0
    with Writer(path) as w:
        for xyz, t in scans:
            end = t[-1]
            while k < len(imu) and imu[k, 0] <= end:
                w.imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
                k += 1
            w.lidar(xyz, t)
            if k < len(imu):  # the first sample past the sweep end makes the frame eligible (Odometry.cpp:65-69)
                w.imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
                k += 1
        while k < len(imu):
            w.imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
