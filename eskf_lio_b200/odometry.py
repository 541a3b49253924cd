"""ctypes binding of include/eskf_host.h (libeskf_host.so): the ROS-free
Odometry + ErrorStateKF host classes whose hot-path calls run on the B200.

No fallback: creation raises without the library or without a CUDA device.
Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build, capi


class OdomConfig(C.Structure):
    """eskf_odom_config: the keys of config/hilti_config.yaml the path reads."""
    _fields_ = [("imu_update_rate", C.c_double), ("bias_a", C.c_double * 3),
                ("bias_g", C.c_double * 3), ("gravity", C.c_double * 3),
                ("accel_noise_density", C.c_double * 3), ("accel_zero_g_offset", C.c_double),
                ("gyro_noise_density", C.c_double), ("gyro_zero_rate_offset", C.c_double),
                ("translation_noise", C.c_double), ("rotation_noise", C.c_double),
                ("lidar_quaternion_xyzw", C.c_double * 4), ("lidar_translation", C.c_double * 3),
                ("map_voxel_size", C.c_double), ("max_points_per_voxel", C.c_uint64),
                ("update_translation_sq_threshold", C.c_double),
                ("update_cosine_threshold", C.c_double), ("remove_enabled", C.c_int32),
                ("remove_distance_threshold", C.c_double), ("remove_period", C.c_double),
                ("preprocess_voxel_size", C.c_double), ("max_iteration", C.c_int32),
                ("neighbor_mode", C.c_int32), ("icp_translation_sq_threshold", C.c_double),
                ("icp_cosine_threshold", C.c_double), ("device_resident", C.c_int32),
                ("map_capacity_hint", C.c_uint64)]


class OdomInfo(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("n_states", C.c_uint64), ("map_voxels", C.c_uint64),
                ("last_removed", C.c_uint64), ("last_iterations", C.c_int32),
                ("last_inserted", C.c_int32), ("stage_avg_ms", C.c_double * 3),
                ("stage_max_ms", C.c_double * 3), ("stage_sum_ms", C.c_double * 3),
                ("device_frame_ms_sum", C.c_double), ("device_frame_ms_last", C.c_double)]


# every symbol include/eskf_host.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["eskf_host_last_error", "eskf_odom_default_config", "eskf_odom_create",
           "eskf_odom_destroy", "eskf_odom_feed_imu", "eskf_odom_feed_lidar",
           "eskf_odom_feed_lidar_cloud", "eskf_odom_context",
           "eskf_odom_spin_once", "eskf_odom_last_pose", "eskf_odom_last_state",
           "eskf_odom_info_get", "eskf_odom_map", "eskf_odom_launch_count",
           "eskf_odom_replay_log", "eskf_log_summary", "eskf_log_writer_open", "eskf_log_writer_imu",
           "eskf_log_writer_lidar", "eskf_log_writer_close"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        capi.lib()  # libeskf_gpu.so first (RTLD_GLOBAL not needed: rpath $ORIGIN)
        path = _build.HOST_LIB_PATH
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.eskf_host_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("eskf_host: " + lib().eskf_host_last_error().decode())


def default_config(**overrides) -> OdomConfig:
    cfg = OdomConfig()
    lib().eskf_odom_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


_dp = C.POINTER(C.c_double)


class Odometry:
    """Odometry (src/Odometry.cpp) + ErrorStateKF (src/ErrorStateKF.cpp), ROS-free."""

    def __init__(self, cfg: OdomConfig | None = None, device: int = 0):
        self.cfg = cfg or default_config()
        self._h = C.c_void_p()
        _check(lib().eskf_odom_create(C.byref(self.cfg), C.c_int(device), C.byref(self._h)))
        self._keep = []  # sweeps whose asynchronous upload may still be in flight

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().eskf_odom_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def feed_imu(self, t, gyro, acc):
        g = np.ascontiguousarray(gyro, dtype=np.float64)
        a = np.ascontiguousarray(acc, dtype=np.float64)
        _check(lib().eskf_odom_feed_imu(self._h, C.c_double(t), g.ctypes.data_as(_dp),
                                        a.ctypes.data_as(_dp)))

    def feed_lidar(self, xyz_f32, point_time):
        """xyz_f32: float32[N,3] (the PointCloud2 wire format); kept alive until consumed."""
        xyz = np.ascontiguousarray(xyz_f32, dtype=np.float32)
        t = np.ascontiguousarray(point_time, dtype=np.float64)
        self._keep.append((xyz, t))
        _check(lib().eskf_odom_feed_lidar(self._h, xyz.ctypes.data_as(C.POINTER(C.c_float)),
                                          t.ctypes.data_as(_dp), C.c_size_t(t.shape[0])))

    def feed_lidar_ptr(self, xyz_ptr, time_ptr, n):
        """Raw pointers (e.g. pinned memory from capi eskf_host_alloc), no copies."""
        _check(lib().eskf_odom_feed_lidar(self._h, C.cast(xyz_ptr, C.POINTER(C.c_float)),
                                          C.cast(time_ptr, _dp), C.c_size_t(n)))

    def feed_lidar_cloud(self, cloud, point_time):
        """cloud: capi.Cloud on self.context() holding the raw sweep (xyz only)."""
        t = np.ascontiguousarray(point_time, dtype=np.float64)
        self._keep.append((cloud, t))
        _check(lib().eskf_odom_feed_lidar_cloud(self._h, cloud._h, t.ctypes.data_as(_dp),
                                                C.c_size_t(t.shape[0])))

    def context(self):
        """The process-wide eskf_ctx of the host classes as a borrowed capi.Context."""
        h = C.c_void_p()
        _check(lib().eskf_odom_context(self._h, C.byref(h)))
        ctx = capi.Context.__new__(capi.Context)
        ctx._h = h
        ctx.device = 0
        ctx.close = lambda: None  # borrowed: never destroyed from Python
        return ctx

    def spin_once(self) -> int:
        c = C.c_int(0)
        _check(lib().eskf_odom_spin_once(self._h, C.byref(c)))
        if c.value:
            self._keep = self._keep[-1:]
        return c.value

    def pose(self):
        T = np.zeros(16)
        _check(lib().eskf_odom_last_pose(self._h, T.ctypes.data_as(_dp)))
        return T.reshape(4, 4)

    def last_state(self, with_P=False):
        s = np.zeros(20)
        P = np.zeros(324) if with_P else None
        _check(lib().eskf_odom_last_state(self._h, s.ctypes.data_as(_dp),
                                          P.ctypes.data_as(_dp) if with_P else None))
        d = {"t": s[0], "p": s[1:4].copy(), "v": s[4:7].copy(), "q": s[7:11].copy(),
             "ba": s[11:14].copy(), "bg": s[14:17].copy(), "g": s[17:20].copy()}
        if with_P:
            d["P"] = P.reshape(18, 18)
        return d

    def info(self) -> OdomInfo:
        out = OdomInfo()
        _check(lib().eskf_odom_info_get(self._h, C.byref(out)))
        return out

    def map_handle(self):
        h = C.c_void_p()
        _check(lib().eskf_odom_map(self._h, C.byref(h)))
        return h

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        _check(lib().eskf_odom_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def replay_log(self, path: str, max_frames: int = 1 << 16):
        """Replay a flat binary sensor log (eskf_lio_b200.sensor_log / SensorLog.hpp) through the
        odometry; returns the poses of the frames that went through."""
        poses = np.zeros((max_frames, 16))
        n = C.c_size_t(0)
        _check(lib().eskf_odom_replay_log(self._h, path.encode(), poses.ctypes.data_as(_dp),
                                          C.c_size_t(max_frames), C.byref(n)))
        return [poses[i].reshape(4, 4).copy() for i in range(min(n.value, max_frames))]


def run_sequence(odom, scans, imu, on_frame=None):
    """Replay a synthetic log the way the two sensor callbacks would deliver it:
    IMU samples in time order, each sweep once its last point has been measured,
    one spin_once per delivery.  scans[i] = (xyz float32/64 [N,3], point_time[N]).
    Returns the per-frame poses (frame 0 = initialisation)."""
    poses = []
    k = 0
    n_imu = imu.shape[0]
    for xyz, t in scans:
        end = t[-1]
        while k < n_imu and imu[k, 0] <= end:
            odom.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
            odom.spin_once()
        odom.feed_lidar(xyz, t)
        done = odom.spin_once()
        while not done and k < n_imu:
            odom.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
            done = odom.spin_once()
        if not done:
            raise RuntimeError("IMU stream ended before the sweep could be processed")
        poses.append(odom.pose())
        if on_frame:
            on_frame(len(poses) - 1, odom)
    return poses
