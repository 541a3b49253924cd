"""ctypes binding of oracle/_ref/libref_shim*.so: the REFERENCE'S OWN sources
(/root/reference/src/{Registration,LocalMap,CloudPreprocessor,Utils,ErrorStateKF,Odometry}.cpp), compiled
where they lie against the API shims in oracle/refshim/include (Eigen, Open3D and yaml-cpp do
not exist in this environment), behind the small C ABI of oracle/refshim/ref_capi.cpp.

TEST INFRASTRUCTURE ONLY.  It pins the ORACLE's restatement of the reference's control flow and
formulas (tests/test_reference_shim.py); the third-party arithmetic underneath is the shims'.
Two builds: ``tree`` sums short inner products as Eigen's unrolled reductions do
(a0 + (a1 + a2)), ``seq`` left to right like the oracle's documented convention.
/root/reference is not on the GPU box: ``available()`` is False there unless the prebuilt
libraries travelled with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
_PATHS = {"tree": os.path.join(_HERE, "_ref", "libref_shim.so"),
          "seq": os.path.join(_HERE, "_ref", "libref_shim_seq.so")}
_libs: dict = {}

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


class Config(C.Structure):
    """RefConfig: the keys of config/hilti_config.yaml the five classes read (same defaults)."""
    _fields_ = [("voxel_map", C.c_double), ("voxel_pre", C.c_double), ("max_points_per_voxel", C.c_uint64),
                ("update_tsq", C.c_double), ("update_cos", C.c_double), ("remove_enabled", C.c_int32),
                ("remove_distance", C.c_double), ("remove_period", C.c_double), ("max_iteration", C.c_int32),
                ("icp_tsq", C.c_double), ("icp_cos", C.c_double), ("lidar_quat_xyzw", C.c_double * 4),
                ("lidar_trans", C.c_double * 3), ("imu_rate", C.c_double), ("bias_a", C.c_double * 3),
                ("bias_g", C.c_double * 3), ("gravity", C.c_double * 3), ("accel_noise_density", C.c_double * 3),
                ("accel_zero_g_offset", C.c_double), ("gyro_noise_density", C.c_double),
                ("gyro_zero_rate_offset", C.c_double), ("translation_noise", C.c_double),
                ("rotation_noise", C.c_double)]


class State(C.Structure):
    _fields_ = [("timestamp", C.c_double), ("position", C.c_double * 3), ("velocity", C.c_double * 3),
                ("attitude_xyzw", C.c_double * 4), ("bias_a", C.c_double * 3), ("bias_g", C.c_double * 3),
                ("gravity", C.c_double * 3)]


def default_config(**overrides) -> Config:
    """config/hilti_config.yaml"""
    c = Config(voxel_map=0.3, voxel_pre=0.3, max_points_per_voxel=1000, update_tsq=1e-2, update_cos=0.985,
               remove_enabled=1, remove_distance=100.0, remove_period=10.0, max_iteration=100, icp_tsq=1e-6,
               icp_cos=0.9999, imu_rate=400.0, accel_zero_g_offset=20.0, gyro_noise_density=0.014,
               gyro_zero_rate_offset=1.0, translation_noise=1e-6, rotation_noise=1e-6)
    c.lidar_quat_xyzw[:] = [0.7071068, -0.7071068, 0.0, 0.0]
    c.lidar_trans[:] = [-0.001, -0.00855, 0.055]
    c.bias_a[:] = [0.06080652138668933, 0.08353074835853214, 0.057072968234636895]
    c.bias_g[:] = [-0.0015351229643790084, -0.0013449146576507546, 0.00030127855524786183]
    c.gravity[:] = [0.01165152782783894, -0.008749296634685332, 9.804989173462031]
    c.accel_noise_density[:] = [105.0, 105.0, 135.0]
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        if isinstance(v, (list, tuple, np.ndarray)):
            getattr(c, k)[:] = list(v)
        else:
            setattr(c, k, v)
    return c


def can_build() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def build(force: bool = False) -> None:
    """make -C oracle ref (needs /root/reference; outputs only into oracle/_ref/)."""
    if not can_build():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present: the reference-on-shim library cannot be built here")
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []) + ["ref"], stdout=subprocess.DEVNULL)


def available() -> bool:
    return all(os.path.exists(p) for p in _PATHS.values()) or can_build()


def lib(kind: str = "tree"):
    if kind not in _libs:
        if can_build():
            build()
        path = _PATHS[kind]
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and {REFERENCE_ROOT} is not present")
        L = C.CDLL(path)
        for name in ("ref_map_create", "ref_eskf_create"):
            getattr(L, name).restype = C.c_void_p
        for name in ("ref_map_size", "ref_correspondences", "ref_gn_step", "ref_eskf_num_states"):
            getattr(L, name).restype = C.c_uint64
        L.ref_preprocess.restype = C.c_longlong
        assert L.ref_tree_redux() == (1 if kind == "tree" else 0)
        _libs[kind] = L
    return _libs[kind]


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def make_states(ts, pos, quat_xyzw):
    n = len(ts)
    arr = (State * max(n, 1))()
    for i in range(n):
        arr[i].timestamp = float(ts[i])
        arr[i].position[:] = [float(v) for v in pos[i]]
        arr[i].attitude_xyzw[:] = [float(v) for v in quat_xyzw[i]]
    return arr, n


class Ref:
    """One build (``tree`` or ``seq``) of the reference-on-shim library."""

    def __init__(self, kind: str = "tree"):
        self.kind = kind
        self.L = lib(kind)

    # ------------------------------------------------------------- Utils
    def skew(self, v):
        o = np.zeros(9)
        self.L.ref_skew(_d(_f64(v)), _d(o))
        return o.reshape(3, 3)

    def rotvec_to_matrix(self, r):
        o = np.zeros(9)
        self.L.ref_rotvec_to_matrix(_d(_f64(r)), _d(o))
        return o.reshape(3, 3)

    def rotation_matrix_to_vector(self, R):
        o = np.zeros(3)
        self.L.ref_rotation_matrix_to_vector(_d(_f64(R)), _d(o))
        return o

    def se3_to_SE3(self, se3):
        o = np.zeros(16)
        self.L.ref_se3_to_SE3(_d(_f64(se3)), _d(o))
        return o.reshape(4, 4)

    def interpolate_SE3(self, s1, s2, t):
        """s = (timestamp, position, quat_xyzw)"""
        a, _ = make_states([s1[0]], [s1[1]], [s1[2]])
        b, _ = make_states([s2[0]], [s2[1]], [s2[2]])
        o = np.zeros(16)
        self.L.ref_interpolate_SE3(C.byref(a[0]), C.byref(b[0]), C.c_double(t), _d(o))
        return o.reshape(4, 4)

    def transform_cloud(self, xyz, cov, T):
        xyz = _f64(xyz, (-1, 3)).copy()
        c = None if cov is None else _f64(cov).reshape(-1, 9).copy()
        self.L.ref_transform_cloud(_d(xyz), None if c is None else _d(c), C.c_size_t(len(xyz)), _d(_f64(T)))
        return xyz, None if c is None else c.reshape(-1, 3, 3)

    def voxel_index(self, xyz, voxel_size):
        xyz = _f64(xyz, (-1, 3))
        out = np.zeros((len(xyz), 3), dtype=np.int32)
        cfg = default_config(voxel_map=voxel_size)
        self.L.ref_voxel_index(C.byref(cfg), _d(xyz), C.c_size_t(len(xyz)), out.ctypes.data_as(_i32p))
        return out

    def preprocess(self, xyz, point_time, T_il, states, voxel_size):
        """CloudPreprocessor::process; rows in the reference's own (hash-map) order."""
        xyz = _f64(xyz, (-1, 3))
        n = len(xyz)
        t = _f64(point_time)
        arr, ns = make_states(*states) if states is not None else make_states([], [], [])
        oxyz = np.zeros((n, 3))
        ocov = np.zeros((n, 9))
        cfg = default_config(voxel_pre=voxel_size)
        m = self.L.ref_preprocess(C.byref(cfg), _d(xyz), _d(t), C.c_size_t(n), _d(_f64(T_il)), arr,
                                  C.c_size_t(ns), _d(oxyz), _d(ocov))
        return oxyz[:m].copy(), ocov[:m].reshape(m, 3, 3).copy()

    def run_odometry(self, scans, imu, state_time, cfg: Config | None = None, **overrides):
        """ESKF_LIO::Odometry::run over a whole log (scans: [(xyz, point_time)], imu: n x 7).
        Returns (poses of every updateLocalMap call, filter state at/before state_time with P,
        map voxel count, number of filter states)."""
        cfg = cfg if cfg is not None else default_config(**overrides)
        xyz = _f64(np.concatenate([x for x, _ in scans]), (-1, 3))
        t = _f64(np.concatenate([tt for _, tt in scans]))
        n_per = np.array([len(tt) for _, tt in scans], dtype=np.uint64)
        imu = _f64(imu, (-1, 7))
        poses = np.zeros((len(scans), 16))
        st = State()
        P = np.zeros((18, 18))
        vox = C.c_uint64(0)
        self.L.ref_odom_run.restype = C.c_longlong
        ns = self.L.ref_odom_run(C.byref(cfg), _d(imu), C.c_size_t(len(imu)), _d(xyz), _d(t),
                                 n_per.ctypes.data_as(_u64p), C.c_size_t(len(scans)), _d(poses),
                                 C.c_double(state_time), C.byref(st), _d(P), C.byref(vox))
        if ns < 0:
            raise RuntimeError("reference Odometry::run did not finish")
        state = {"timestamp": st.timestamp, "position": np.array(st.position), "velocity": np.array(st.velocity),
                 "attitude_xyzw": np.array(st.attitude_xyzw), "bias_a": np.array(st.bias_a),
                 "bias_g": np.array(st.bias_g), "gravity": np.array(st.gravity), "P": P}
        return poses.reshape(-1, 4, 4), state, int(vox.value), int(ns)

    def Map(self, cfg: Config | None = None, **overrides):
        return RefMap(self, cfg if cfg is not None else default_config(**overrides))

    def Eskf(self, cfg: Config | None = None, **overrides):
        return RefEskf(self, cfg if cfg is not None else default_config(**overrides))


class RefMap:
    """ESKF_LIO::LocalMap + ESKF_LIO::ICP of the reference."""

    def __init__(self, ref: Ref, cfg: Config):
        self.L = ref.L
        self.cfg = cfg
        self._h = C.c_void_p(self.L.ref_map_create(C.byref(cfg)))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_map_destroy(self._h)
            self._h = None

    def update(self, xyz, cov, T, initialize=False):
        xyz = _f64(xyz, (-1, 3)).copy()
        cov = _f64(cov).reshape(-1, 9).copy()
        self.L.ref_map_update(self._h, _d(xyz), _d(cov), C.c_size_t(len(xyz)), _d(_f64(T)), C.c_int(int(initialize)))
        return xyz, cov.reshape(-1, 3, 3)

    def size(self) -> int:
        return int(self.L.ref_map_size(self._h))

    def export(self):
        """sorted by (kx, ky, kz) like the oracle's export"""
        n = self.size()
        keys = np.zeros((n, 3), dtype=np.int32)
        count = np.zeros(n, dtype=np.uint64)
        mean = np.zeros((n, 3))
        cov = np.zeros((n, 9))
        self.L.ref_map_export(self._h, keys.ctypes.data_as(_i32p), count.ctypes.data_as(_u64p), _d(mean), _d(cov))
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        return keys[order], count[order], mean[order], cov[order].reshape(n, 3, 3)

    def needs_map_update(self, prev, cur) -> bool:
        return bool(self.L.ref_needs_map_update(self._h, _d(_f64(prev)), _d(_f64(cur))))

    def correspondences(self, xyz, cov):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        n = len(xyz)
        sp, sc, mp, mc = np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 3)), np.zeros((n, 9))
        m = int(self.L.ref_correspondences(self._h, _d(xyz), _d(cov), C.c_size_t(n), _d(sp), _d(sc), _d(mp), _d(mc)))
        return sp[:m], sc[:m].reshape(m, 3, 3), mp[:m], mc[:m].reshape(m, 3, 3)

    def jtj_jtr(self, p, mu, Cm):
        H, b = np.zeros(36), np.zeros(6)
        self.L.ref_jtj_jtr(self._h, _d(_f64(p)), _d(_f64(mu)), _d(_f64(Cm)), _d(H), _d(b))
        return H.reshape(6, 6), b

    def gn_step(self, xyz, cov):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        T = np.zeros(16)
        nc = int(self.L.ref_gn_step(self._h, _d(xyz), _d(cov), C.c_size_t(len(xyz)), _d(T)))
        return T.reshape(4, 4), nc

    def convergence_check(self, T) -> bool:
        return bool(self.L.ref_convergence_check(self._h, _d(_f64(T))))

    def align(self, xyz, cov, guess):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        T = np.zeros(16)
        conv = self.L.ref_align(self._h, _d(xyz), _d(cov), C.c_size_t(len(xyz)), _d(_f64(guess)), _d(T))
        return T.reshape(4, 4), bool(conv)


class RefEskf:
    """ESKF_LIO::ErrorStateKF of the reference."""

    def __init__(self, ref: Ref, cfg: Config):
        self.L = ref.L
        self._h = C.c_void_p(self.L.ref_eskf_create(C.byref(cfg)))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_eskf_destroy(self._h)
            self._h = None

    def feed_imu(self, t, gyro, acc):
        self.L.ref_eskf_feed_imu(self._h, C.c_double(t), _d(_f64(gyro)), _d(_f64(acc)))

    def initialize(self, lidar_end_time):
        self.L.ref_eskf_initialize(self._h, C.c_double(lidar_end_time))

    def process(self, t, gyro, acc):
        self.L.ref_eskf_process(self._h, C.c_double(t), _d(_f64(gyro)), _d(_f64(acc)))

    def update(self, rmap: RefMap, xyz, cov, lidar_end_time):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        T = np.zeros(16)
        self.L.ref_eskf_update(self._h, rmap._h, _d(xyz), _d(cov), C.c_size_t(len(xyz)), C.c_double(lidar_end_time), _d(T))
        return T.reshape(4, 4)

    def num_states(self) -> int:
        return int(self.L.ref_eskf_num_states(self._h))

    def state(self, index=-1, with_P=False):
        s = State()
        P = np.zeros((18, 18)) if with_P else None
        self.L.ref_eskf_state(self._h, C.c_longlong(index), C.byref(s), _d(P) if with_P else None)
        out = {"timestamp": s.timestamp, "position": np.array(s.position), "velocity": np.array(s.velocity),
               "attitude_xyzw": np.array(s.attitude_xyzw), "bias_a": np.array(s.bias_a),
               "bias_g": np.array(s.bias_g), "gravity": np.array(s.gravity)}
        if with_P:
            out["P"] = P
        return out
