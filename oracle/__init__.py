"""ctypes binding of the CPU ORACLE (oracle/eskf_oracle.cpp).

TEST INFRASTRUCTURE ONLY — may be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs, never by the product
package ``eskf_lio_b200``.  The reference has no tests and cannot be built as shipped;
see eskf_oracle.h for how the oracle is pinned instead (incl. oracle/ref.py: the reference's
own sources compiled against API shims).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("eskf_oracle.cpp", "odom_oracle.cpp", "eskf_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class IcpParams(C.Structure):
    _fields_ = [("max_iteration", C.c_int32), ("neighbor_mode", C.c_int32),
                ("translation_sq_threshold", C.c_double), ("cosine_threshold", C.c_double)]


class AlignInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32)]


class State(C.Structure):
    _fields_ = [("timestamp", C.c_double), ("position", C.c_double * 3),
                ("attitude_xyzw", C.c_double * 4)]


def lib():
    global _lib
    if _lib is None:
        # never build implicitly on import when the .so is already there (the
        # GPU box has no need to recompile); build() is explicit.
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_double, C.c_uint64]
        L.orc_map_size.restype = C.c_uint64
        L.orc_map_evict.restype = C.c_uint64
        L.orc_linearize.restype = C.c_size_t
        L.orc_downsample_cov.restype = C.c_size_t
        L.orc_preprocess.restype = C.c_long
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ------------------------------------------------------------------ Utils
def skew(v):
    out = np.zeros(9)
    lib().orc_skew(_d(_f64(v)), _d(out))
    return out.reshape(3, 3)


def compute_J(r):
    out = np.zeros(9)
    lib().orc_compute_J(_d(_f64(r)), _d(out))
    return out.reshape(3, 3)


def rotvec_to_matrix(r):
    out = np.zeros(9)
    lib().orc_rotvec_to_matrix(_d(_f64(r)), _d(out))
    return out.reshape(3, 3)


def se3_to_SE3(se3):
    out = np.zeros(16)
    lib().orc_se3_to_SE3(_d(_f64(se3)), _d(out))
    return out.reshape(4, 4)


def quat_to_matrix(q_xyzw):
    out = np.zeros(9)
    lib().orc_quat_to_matrix(_d(_f64(q_xyzw)), _d(out))
    return out.reshape(3, 3)


def make_states(ts, pos, quat_xyzw):
    n = len(ts)
    arr = (State * max(n, 1))()
    for i in range(n):
        arr[i].timestamp = float(ts[i])
        for j in range(3):
            arr[i].position[j] = float(pos[i][j])
        for j in range(4):
            arr[i].attitude_xyzw[j] = float(quat_xyzw[i][j])
    return arr, n


def interpolate_SE3(s1, s2, t):
    arr, _ = make_states([s1[0], s2[0]], [s1[1], s2[1]], [s1[2], s2[2]])
    out = np.zeros(16)
    lib().orc_interpolate_SE3(C.byref(arr[0]), C.byref(arr[1]), C.c_double(t), _d(out))
    return out.reshape(4, 4)


def transform_cloud(xyz, cov, T):
    xyz = _f64(xyz).copy()
    cov_c = None if cov is None else _f64(cov).reshape(-1, 9).copy()
    lib().orc_transform_cloud(_d(xyz), None if cov_c is None else _d(cov_c),
                              C.c_size_t(xyz.shape[0]), _d(_f64(T)))
    return xyz, (None if cov_c is None else cov_c.reshape(-1, 3, 3))


def voxel_index(xyz, voxel_size):
    xyz = _f64(xyz, (-1, 3))
    out = np.zeros((xyz.shape[0], 3), dtype=np.int32)
    lib().orc_voxel_index(_d(xyz), C.c_size_t(xyz.shape[0]), C.c_double(voxel_size),
                          out.ctypes.data_as(_i32p))
    return out


# ----------------------------------------------------------- Registration
def jtj_jtr(p, mu, Cm):
    H = np.zeros(36)
    b = np.zeros(6)
    lib().orc_jtj_jtr(_d(_f64(p)), _d(_f64(mu)), _d(_f64(Cm)), _d(H), _d(b))
    return H.reshape(6, 6), b


def ldlt_solve6(H, b):
    x = np.zeros(6)
    lib().orc_ldlt_solve6(_d(_f64(H)), _d(_f64(b)), _d(x))
    return x


def convergence_check(T, trans_sq_thr, cos_thr) -> bool:
    return bool(lib().orc_convergence_check(_d(_f64(T)), C.c_double(trans_sq_thr),
                                            C.c_double(cos_thr)))


def needs_map_update(prev, cur, trans_sq_thr, cos_thr) -> bool:
    return bool(lib().orc_needs_map_update(_d(_f64(prev)), _d(_f64(cur)),
                                           C.c_double(trans_sq_thr), C.c_double(cos_thr)))


class Map:
    """LocalMap restatement (src/LocalMap.cpp)."""

    def __init__(self, voxel_size: float, max_points_per_voxel: int = 1000):
        self._h = C.c_void_p(lib().orc_map_create(C.c_double(voxel_size),
                                                  C.c_uint64(max_points_per_voxel)))
        self.voxel_size = voxel_size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_map_destroy(self._h)
            self._h = None

    def set_update_params(self, trans_sq_thr=1e-2, cos_thr=0.985, remove_enabled=True,
                          distance_thr=100.0, remove_period=10.0):
        lib().orc_map_set_update_params(self._h, C.c_double(trans_sq_thr), C.c_double(cos_thr),
                                        C.c_int(int(remove_enabled)), C.c_double(distance_thr),
                                        C.c_double(remove_period))

    def update(self, xyz, cov, T, initialize=False, now=0.0):
        """updateLocalMap; returns (inserted, removed, xyz_world, cov_world)."""
        xyz = _f64(xyz, (-1, 3)).copy()
        cov = _f64(cov).reshape(-1, 9).copy()
        removed = C.c_uint64(0)
        ins = lib().orc_map_update(self._h, _d(xyz), _d(cov), C.c_size_t(xyz.shape[0]),
                                   _d(_f64(T)), C.c_int(int(initialize)), C.c_double(now),
                                   C.byref(removed))
        return bool(ins), int(removed.value), xyz, cov.reshape(-1, 3, 3)

    def insert(self, xyz, cov):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        lib().orc_map_insert(self._h, _d(xyz), _d(cov), C.c_size_t(xyz.shape[0]))

    def evict(self, pos, distance_thr) -> int:
        return int(lib().orc_map_evict(self._h, _d(_f64(pos)), C.c_double(distance_thr)))

    def size(self) -> int:
        return int(lib().orc_map_size(self._h))

    def export(self):
        n = self.size()
        keys = np.zeros((n, 3), dtype=np.int32)
        count = np.zeros(n, dtype=np.uint64)
        mean = np.zeros((n, 3))
        cov = np.zeros((n, 9))
        lib().orc_map_export(self._h, keys.ctypes.data_as(_i32p), count.ctypes.data_as(_u64p),
                             _d(mean), _d(cov))
        return keys, count, mean, cov.reshape(n, 3, 3)

    def query(self, xyz):
        xyz = _f64(xyz, (-1, 3))
        n = xyz.shape[0]
        keys = np.zeros((n, 3), dtype=np.int32)
        hit = np.zeros(n, dtype=np.uint8)
        count = np.zeros(n, dtype=np.uint64)
        mean = np.zeros((n, 3))
        cov = np.zeros((n, 9))
        lib().orc_map_query(self._h, _d(xyz), C.c_size_t(n), keys.ctypes.data_as(_i32p),
                            hit.ctypes.data_as(_u8p), count.ctypes.data_as(_u64p), _d(mean),
                            _d(cov))
        return keys, hit.astype(bool), count, mean, cov.reshape(n, 3, 3)

    def linearize(self, xyz, cov, neighbor_mode=1):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        n = xyz.shape[0]
        nn = 7 if neighbor_mode == 7 else 1
        H = np.zeros(36)
        b = np.zeros(6)
        hit = np.zeros(n * nn, dtype=np.uint8)
        nc = lib().orc_linearize(self._h, _d(xyz), _d(cov), C.c_size_t(n), C.c_int(neighbor_mode),
                                 _d(H), _d(b), hit.ctypes.data_as(_u8p))
        return H.reshape(6, 6), b, hit.reshape(n, nn).astype(bool), int(nc)

    def align(self, xyz, cov, guess, max_iteration=100, translation_sq_threshold=1e-6,
              cosine_threshold=0.9999, neighbor_mode=1):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        info = AlignInfo()
        T = np.zeros(16)
        tH = np.zeros((max_iteration, 36))
        tb = np.zeros((max_iteration, 6))
        tn = np.zeros(max_iteration, dtype=np.uint64)
        ts = np.zeros((max_iteration, 16))
        lib().orc_align(self._h, _d(xyz), _d(cov), C.c_size_t(xyz.shape[0]), _d(_f64(guess)),
                        C.byref(prm), _d(T), C.byref(info), _d(tH), _d(tb),
                        tn.ctypes.data_as(_u64p), _d(ts))
        it = info.iterations
        return {"T": T.reshape(4, 4), "iterations": it, "converged": bool(info.converged),
                "H": tH[:it].reshape(it, 6, 6), "b": tb[:it], "ncorr": tn[:it].astype(np.int64),
                "step": ts[:it].reshape(it, 4, 4)}


# ------------------------------------------------------ CloudPreprocessor
def knn(xyz, queries, k=30, bruteforce=False):
    xyz = _f64(xyz, (-1, 3))
    q = _f64(queries, (-1, 3))
    idx = np.zeros((q.shape[0], k), dtype=np.int32)
    d2 = np.zeros((q.shape[0], k))
    fn = lib().orc_knn_bruteforce if bruteforce else lib().orc_knn
    fn(_d(xyz), C.c_size_t(xyz.shape[0]), _d(q), C.c_size_t(q.shape[0]), C.c_int(k),
       idx.ctypes.data_as(_i32p), _d(d2))
    return idx, d2


def cov_from_indices(xyz, idx):
    xyz = _f64(xyz, (-1, 3))
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(9)
    lib().orc_cov_from_indices(_d(xyz), idx.ctypes.data_as(_i32p), C.c_int(idx.shape[0]), _d(out))
    return out.reshape(3, 3)


def regularize_cov(Cm):
    out = np.zeros(9)
    lib().orc_regularize_cov(_d(_f64(Cm)), _d(out))
    return out.reshape(3, 3)


def downsample_cov(xyz, voxel_size):
    xyz = _f64(xyz, (-1, 3))
    n = xyz.shape[0]
    oxyz = np.zeros((n, 3))
    ocov = np.zeros((n, 9))
    osrc = np.zeros(n, dtype=np.uint32)
    m = lib().orc_downsample_cov(_d(xyz), C.c_size_t(n), C.c_double(voxel_size), _d(oxyz),
                                 _d(ocov), osrc.ctypes.data_as(_u32p))
    return oxyz[:m].copy(), ocov[:m].reshape(m, 3, 3).copy(), osrc[:m].copy()


def deskew(xyz, point_time, states):
    """states = (ts, pos, quat_xyzw) arrays."""
    xyz = _f64(xyz, (-1, 3)).copy()
    t = _f64(point_time)
    arr, ns = make_states(*states)
    rc = lib().orc_deskew(_d(xyz), _d(t), C.c_size_t(xyz.shape[0]), arr, C.c_size_t(ns))
    if rc != 0:
        raise RuntimeError("deskew: no state at or before the scan end time")
    return xyz


def preprocess(xyz, point_time, T_il, states, voxel_size, min_range=0.0, max_range=0.0):
    """CloudPreprocessor::process (src/CloudPreprocessor.cpp:10-23).  min_range / max_range: the range
    crop BASELINE.json's north_star names and the reference lacks, DEFINED here: a point survives iff
    min^2 <= ((x*x + y*y) + z*z) <= max^2 in the LiDAR frame (max 0 = unbounded); T_il (:16) and the
    deskew (:17-19) act on every point of the sweep, the cropped ones are erased ahead of
    voxelDownsampleAndEstimateCovariances (:22); returned source indices index the uncropped sweep."""
    if min_range > 0.0 or max_range > 0.0:
        raw = _f64(xyz, (-1, 3))
        r2 = (raw[:, 0] * raw[:, 0] + raw[:, 1] * raw[:, 1]) + raw[:, 2] * raw[:, 2]
        keep = r2 >= min_range * min_range
        if max_range > 0.0:
            keep &= r2 <= max_range * max_range
        p, _ = transform_cloud(raw, None, T_il)
        if states is not None and len(states[0]) > 0:
            p = deskew(p, point_time, states)
        idx = np.nonzero(keep)[0]
        if len(idx) == 0:
            return np.zeros((0, 3)), np.zeros((0, 3, 3)), np.zeros(0, dtype=np.uint32)
        op, oc, osrc = downsample_cov(p[idx], voxel_size)
        return op, oc, idx[osrc].astype(np.uint32)
    xyz = _f64(xyz, (-1, 3)).copy()
    n = xyz.shape[0]
    t = _f64(point_time if point_time is not None else np.zeros(n))
    if states is None:
        arr, ns = make_states([], [], [])
    else:
        arr, ns = make_states(*states)
    oxyz = np.zeros((n, 3))
    ocov = np.zeros((n, 9))
    osrc = np.zeros(n, dtype=np.uint32)
    m = lib().orc_preprocess(_d(xyz), _d(t), C.c_size_t(n), _d(_f64(T_il)), arr, C.c_size_t(ns),
                             C.c_double(voxel_size), _d(oxyz), _d(ocov),
                             osrc.ctypes.data_as(_u32p))
    if m < 0:
        raise RuntimeError("preprocess: deskew failed")
    return oxyz[:m].copy(), ocov[:m].reshape(m, 3, 3).copy(), osrc[:m].copy()


# ------------------------------------------- ErrorStateKF + Odometry::run
class OdomConfig(C.Structure):
    """orc_odom_config: the keys of config/hilti_config.yaml the path reads."""
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
                ("icp_cosine_threshold", C.c_double)]


class OdomInfo(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("n_states", C.c_uint64), ("map_voxels", C.c_uint64),
                ("last_kept", C.c_uint64), ("last_removed", C.c_uint64),
                ("last_iterations", C.c_int32), ("last_inserted", C.c_int32),
                ("stage_avg_ms", C.c_double * 3), ("stage_max_ms", C.c_double * 3)]


def odom_default_config(**overrides) -> OdomConfig:
    cfg = OdomConfig()
    lib().orc_odom_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class Odometry:
    """ErrorStateKF (src/ErrorStateKF.cpp) driven in the call order of
    Odometry::run (src/Odometry.cpp:16-98), on the CPU oracle's hot path."""

    def __init__(self, cfg: OdomConfig | None = None):
        L = lib()
        L.orc_odom_create.restype = C.c_void_p
        L.orc_odom_map.restype = C.c_void_p
        self.cfg = cfg or odom_default_config()
        self._h = C.c_void_p(L.orc_odom_create(C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_odom_destroy(self._h)
            self._h = None

    def feed_imu(self, t, gyro, acc):
        lib().orc_odom_feed_imu(self._h, C.c_double(t), _d(_f64(gyro)), _d(_f64(acc)))

    def feed_lidar(self, xyz, point_time):
        xyz = _f64(xyz, (-1, 3))
        t = _f64(point_time)
        lib().orc_odom_feed_lidar(self._h, _d(xyz), _d(t), C.c_size_t(xyz.shape[0]),
                                  C.c_double(t[0]), C.c_double(t[-1]))

    def spin_once(self) -> int:
        rc = lib().orc_odom_spin_once(self._h)
        if rc < 0:
            raise RuntimeError("odometry: deskew failed")
        return rc

    def kf_process(self, t, gyro, acc):
        lib().orc_kf_process(self._h, C.c_double(t), _d(_f64(gyro)), _d(_f64(acc)))

    def kf_update_with_observation(self, lidar_end, obs):
        guess = np.zeros(16)
        T = np.zeros(16)
        lib().orc_kf_update_with_observation(self._h, C.c_double(lidar_end), _d(_f64(obs)),
                                             _d(guess), _d(T))
        return guess.reshape(4, 4), T.reshape(4, 4)

    def pose(self):
        T = np.zeros(16)
        lib().orc_odom_last_pose(self._h, _d(T))
        return T.reshape(4, 4)

    def info(self) -> OdomInfo:
        out = OdomInfo()
        lib().orc_odom_info(self._h, C.byref(out))
        return out

    def last_state(self, with_P=False):
        s = np.zeros(20)
        P = np.zeros(324) if with_P else None
        lib().orc_odom_last_state(self._h, _d(s), _d(P) if with_P else None)
        d = {"t": s[0], "p": s[1:4].copy(), "v": s[4:7].copy(), "q": s[7:11].copy(),
             "ba": s[11:14].copy(), "bg": s[14:17].copy(), "g": s[17:20].copy()}
        if with_P:
            d["P"] = P.reshape(18, 18)
        return d


def rotation_matrix_to_vector(R):
    out = np.zeros(3)
    lib().orc_rotation_matrix_to_vector(_d(_f64(R)), _d(out))
    return out


def run_sequence(odom, scans, imu, on_frame=None):
    """Feed a synthetic sequence the way the two ROS callbacks would: IMU samples
    in time order, each sweep when its last point has been measured; one
    spin_once per sample.  Returns the list of per-frame poses (frame 0 = init)."""
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
