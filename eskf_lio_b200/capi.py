"""ctypes binding of include/eskf_gpu.h (libeskf_gpu.so, built in-tree).

There is no fallback of any kind: if the library is missing or no B200 is
present, calls raise.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

OK = 0
STATUS_NAMES = {0: "OK", 1: "ERR_CUDA", 2: "ERR_INVALID", 3: "ERR_NO_DEVICE", 4: "ERR_CAPACITY",
                5: "ERR_RANGE", 6: "ERR_INTERNAL"}


class EskfError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"eskf_gpu {STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class IcpParams(C.Structure):
    _fields_ = [("max_iteration", C.c_int32), ("neighbor_mode", C.c_int32),
                ("translation_sq_threshold", C.c_double), ("cosine_threshold", C.c_double)]


class AlignInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("n_corr_last", C.c_uint64),
                ("trace_H", C.POINTER(C.c_double)), ("trace_b", C.POINTER(C.c_double)),
                ("trace_ncorr", C.POINTER(C.c_uint64)), ("trace_step", C.POINTER(C.c_double))]


class State(C.Structure):
    _fields_ = [("timestamp", C.c_double), ("position", C.c_double * 3),
                ("attitude_xyzw", C.c_double * 4)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

# every symbol include/eskf_gpu.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "eskf_abi_version", "eskf_last_error", "eskf_device_count", "eskf_host_alloc", "eskf_host_free",
    "eskf_ctx_create", "eskf_ctx_destroy", "eskf_ctx_sync", "eskf_ctx_stream",
    "eskf_ctx_launch_count", "eskf_ctx_set_option", "eskf_ctx_get_option", "eskf_ctx_timer_start", "eskf_ctx_timer_stop",
    "eskf_cloud_create", "eskf_cloud_destroy", "eskf_cloud_upload", "eskf_cloud_upload_f32",
    "eskf_cloud_download", "eskf_cloud_size", "eskf_cloud_transform", "eskf_cloud_copy",
    "eskf_map_create", "eskf_map_destroy", "eskf_map_insert", "eskf_map_insert_cloud",
    "eskf_map_evict", "eskf_map_size", "eskf_map_capacity", "eskf_map_compact", "eskf_map_query", "eskf_map_export",
    "eskf_ctx_set_range_crop", "eskf_stamps_sorted", "eskf_preprocess", "eskf_preprocess_cloud", "eskf_downsample_cov",
    "eskf_align", "eskf_align_cloud", "eskf_align_cloud_begin", "eskf_align_end", "eskf_align_batch",
    "eskf_linearize",
    "eskf_align_cloud_fixed",
    "eskf_align_cloud_sharded",
    "eskf_comm_create", "eskf_comm_destroy", "eskf_comm_local_handle", "eskf_comm_connect",
    "eskf_align_cloud_p2p",
]

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def lib():
    """Load libeskf_gpu.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("ESKF_GPU_LIB") or _build.LIB_PATH  # env: kernel-variant experiments
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the GPU path has no CPU fallback)")
        L = C.CDLL(path)
        L.eskf_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)  # AttributeError if the ABI is incomplete
        _lib = L
    return _lib


def stamps_sorted(point_time) -> bool:
    """eskf_stamps_sorted: are the per-point stamps non-decreasing (host-only, no GPU needed)."""
    t = _f64(point_time)
    f = lib().eskf_stamps_sorted
    f.restype = C.c_int
    return bool(f(_d(t), C.c_size_t(t.shape[0])))


def check(status: int):
    if status != OK:
        raise EskfError(status, lib().eskf_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    st = lib().eskf_device_count(C.byref(n))
    return n.value if st == OK else 0


_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _opt(a, ptr_type):
    return None if a is None else a.ctypes.data_as(ptr_type)


def make_states(states):
    """states: None or (timestamps[n], positions[n,3], quats_xyzw[n,4])."""
    if states is None:
        return None, 0
    ts, pos, quat = states
    n = len(ts)
    arr = (State * max(n, 1))()
    for i in range(n):
        arr[i].timestamp = float(ts[i])
        for j in range(3):
            arr[i].position[j] = float(pos[i][j])
        for j in range(4):
            arr[i].attitude_xyzw[j] = float(quat[i][j])
    return arr, n


class Context:
    """eskf_ctx: one device, one stream, one caller thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        check(lib().eskf_ctx_create(C.c_int(device), C.c_void_p(stream) if stream else None,
                                    C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().eskf_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().eskf_ctx_sync(self._h))

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().eskf_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        check(lib().eskf_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def set_option(self, name: str, value: int):
        check(lib().eskf_ctx_set_option(self._h, name.encode(), C.c_int64(int(value))))

    def get_option(self, name: str) -> int:
        v = C.c_int64(0)
        check(lib().eskf_ctx_get_option(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    def set_range_crop(self, min_range: float = 0.0, max_range: float = 0.0):
        """Range crop of the preprocessor (LiDAR frame; max_range 0 = unbounded; (0, 0) = off)."""
        check(lib().eskf_ctx_set_range_crop(self._h, C.c_double(min_range), C.c_double(max_range)))

    def align_end(self):
        info, bufs = _make_info(1, False)
        T = np.zeros(16)
        check(lib().eskf_align_end(self._h, _d(T), C.byref(info)))
        return _info_dict(T, info, bufs)

    def timer_start(self):
        check(lib().eskf_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(lib().eskf_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # ---- host-buffer entry points -------------------------------------
    def preprocess(self, xyz, point_time, T_il, states, voxel_size):
        xyz = _f64(xyz, (-1, 3))
        n = xyz.shape[0]
        t = None if point_time is None else _f64(point_time)
        arr, ns = make_states(states)
        oxyz = np.empty((n, 3))
        ocov = np.empty((n, 9))
        osrc = np.empty(n, dtype=np.uint32)
        m = C.c_size_t(0)
        check(lib().eskf_preprocess(self._h, _d(xyz), None if t is None else _d(t), C.c_size_t(n),
                                    None if T_il is None else _d(_f64(T_il)), arr, C.c_size_t(ns),
                                    C.c_double(voxel_size), C.byref(m), _d(oxyz), _d(ocov),
                                    osrc.ctypes.data_as(C.POINTER(C.c_uint32))))
        k = m.value
        return oxyz[:k].copy(), ocov[:k].reshape(k, 3, 3).copy(), osrc[:k].copy()

    def downsample_cov(self, xyz, voxel_size):
        return self.preprocess(xyz, None, None, None, voxel_size)

    def align(self, gmap, xyz, cov, guess, max_iteration=100, translation_sq_threshold=1e-6,
              cosine_threshold=0.9999, neighbor_mode=1, trace=True):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        info, bufs = _make_info(max_iteration, trace)
        T = np.zeros(16)
        check(lib().eskf_align(self._h, gmap._h, _d(xyz), _d(cov), C.c_size_t(xyz.shape[0]),
                               _d(_f64(guess)), C.byref(prm), _d(T), C.byref(info)))
        return _info_dict(T, info, bufs)

    def linearize(self, gmap, xyz, cov, T=None, neighbor_mode=1, fp64_math=False):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        n = xyz.shape[0]
        nn = 7 if neighbor_mode == 7 else 1
        H = np.zeros(36)
        b = np.zeros(6)
        hit = np.zeros(n * nn, dtype=np.uint8)
        nc = C.c_uint64(0)
        T = np.eye(4) if T is None else T
        check(lib().eskf_linearize(self._h, gmap._h, _d(xyz), _d(cov), C.c_size_t(n), _d(_f64(T)),
                                   C.c_int(neighbor_mode), C.c_int(int(fp64_math)), _d(H), _d(b),
                                   hit.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nc)))
        return H.reshape(6, 6), b, hit.reshape(n, nn).astype(bool), int(nc.value)


def _make_info(max_iteration, trace):
    info = AlignInfo()
    bufs = None
    if trace:
        bufs = (np.zeros((max_iteration, 36)), np.zeros((max_iteration, 6)),
                np.zeros(max_iteration, dtype=np.uint64), np.zeros((max_iteration, 16)))
        info.trace_H = _d(bufs[0])
        info.trace_b = _d(bufs[1])
        info.trace_ncorr = bufs[2].ctypes.data_as(C.POINTER(C.c_uint64))
        info.trace_step = _d(bufs[3])
    return info, bufs


def _info_dict(T, info, bufs):
    it = info.iterations
    out = {"T": T.reshape(4, 4).copy(), "iterations": it, "converged": bool(info.converged),
           "n_corr_last": int(info.n_corr_last)}
    if bufs is not None:
        out.update({"H": bufs[0][:it].reshape(it, 6, 6).copy(), "b": bufs[1][:it].copy(),
                    "ncorr": bufs[2][:it].astype(np.int64), "step": bufs[3][:it].reshape(it, 4, 4).copy()})
    return out


class Cloud:
    """eskf_cloud: device-resident PointCloud (points_ + covariances_)."""

    def __init__(self, ctx: Context, capacity: int = 64):
        self.ctx = ctx
        self._h = C.c_void_p()
        check(lib().eskf_cloud_create(ctx._h, C.c_size_t(capacity), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self.ctx._h.value:  # a closed context already released the device
                lib().eskf_cloud_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, xyz, cov=None):
        xyz = _f64(xyz, (-1, 3))
        cov_c = None if cov is None else _f64(cov).reshape(-1, 9)
        check(lib().eskf_cloud_upload(self._h, _d(xyz), None if cov_c is None else _d(cov_c),
                                      C.c_size_t(xyz.shape[0])))
        return self

    def upload_ptr(self, xyz_ptr: int, cov_ptr: int | None, n: int):
        """upload from raw host pointers (e.g. pinned buffers); asynchronous."""
        check(lib().eskf_cloud_upload(self._h, C.cast(xyz_ptr, _dp),
                                      C.cast(cov_ptr, _dp) if cov_ptr else None, C.c_size_t(n)))
        return self

    def upload_f32(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        check(lib().eskf_cloud_upload_f32(self._h, xyz.ctypes.data_as(C.POINTER(C.c_float)),
                                          C.c_size_t(xyz.shape[0])))
        return self

    def size(self) -> int:
        n = C.c_size_t(0)
        check(lib().eskf_cloud_size(self._h, C.byref(n)))
        return n.value

    def download(self, want_cov=True, want_src=False):
        n = self.size()
        xyz = np.empty((n, 3))
        cov = np.empty((n, 9)) if want_cov else None
        src = np.empty(n, dtype=np.uint32) if want_src else None
        m = C.c_size_t(0)
        check(lib().eskf_cloud_download(self._h, _d(xyz), _opt(cov, _dp),
                                        _opt(src, C.POINTER(C.c_uint32)), C.c_size_t(n), C.byref(m)))
        return xyz, (None if cov is None else cov.reshape(n, 3, 3)), src

    def transform(self, T):
        check(lib().eskf_cloud_transform(self._h, _d(_f64(T))))

    def copy_from(self, other: "Cloud"):
        check(lib().eskf_cloud_copy(self._h, other._h))
        return self

    def preprocess_into(self, out: "Cloud", point_time, T_il, states, voxel_size):
        t = None if point_time is None else _f64(point_time)
        arr, ns = make_states(states)
        check(lib().eskf_preprocess_cloud(self.ctx._h, self._h, None if t is None else _d(t),
                                          None if T_il is None else _d(_f64(T_il)), arr,
                                          C.c_size_t(ns), C.c_double(voxel_size), out._h))
        return out


class Comm:
    """eskf_comm: peer-mapped H/b mailboxes of a registration sharded over ranks.

    handles are exchanged by the caller (``all_gather`` takes a list of this rank's
    64-byte handle and returns every rank's, in rank order)."""

    HANDLE_BYTES = 64

    def __init__(self, ctx: "Context", rank: int, world: int, all_gather=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        check(lib().eskf_comm_create(ctx._h, C.c_int(rank), C.c_int(world), C.byref(self._h)))
        if world > 1:
            if all_gather is None:
                raise ValueError("world > 1 needs an all_gather(bytes) -> list[bytes] callable")
            mine = (C.c_ubyte * self.HANDLE_BYTES)()
            check(lib().eskf_comm_local_handle(self._h, mine))
            handles = all_gather(bytes(mine))
            if len(handles) != world or any(len(h) != self.HANDLE_BYTES for h in handles):
                raise ValueError("all_gather must return one 64-byte handle per rank")
            blob = b"".join(handles)
            check(lib().eskf_comm_connect(self._h, blob))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self.ctx._h.value:
                lib().eskf_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Map:
    """eskf_map: the LocalMap voxel hash table in HBM."""

    def __init__(self, ctx: Context, voxel_size: float, max_points_per_voxel: int = 1000,
                 capacity_hint: int = 1 << 16):
        self.ctx = ctx
        self.voxel_size = voxel_size
        self._h = C.c_void_p()
        check(lib().eskf_map_create(ctx._h, C.c_double(voxel_size), C.c_uint32(max_points_per_voxel),
                                    C.c_uint64(capacity_hint), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self.ctx._h.value:  # a closed context already released the device
                lib().eskf_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert(self, xyz, cov, T):
        xyz = _f64(xyz, (-1, 3))
        cov = _f64(cov).reshape(-1, 9)
        check(lib().eskf_map_insert(self._h, _d(xyz), _d(cov), C.c_size_t(xyz.shape[0]), _d(_f64(T))))

    def insert_cloud(self, cloud: Cloud, T):
        check(lib().eskf_map_insert_cloud(self._h, cloud._h, _d(_f64(T))))

    def evict(self, pos, dist_thresh) -> int:
        r = C.c_uint64(0)
        check(lib().eskf_map_evict(self._h, _d(_f64(pos)), C.c_double(dist_thresh), C.byref(r)))
        return r.value

    def size(self) -> int:
        n = C.c_uint64(0)
        check(lib().eskf_map_size(self._h, C.byref(n)))
        return n.value

    def capacity(self) -> int:
        n = C.c_uint64(0)
        check(lib().eskf_map_capacity(self._h, C.byref(n)))
        return n.value

    def compact(self):
        """Re-hash into a table of 2 x size() slots (load factor 1/2, smallest tag array)."""
        check(lib().eskf_map_compact(self._h))

    def query(self, xyz):
        xyz = _f64(xyz, (-1, 3))
        n = xyz.shape[0]
        keys = np.zeros((n, 3), dtype=np.int32)
        hit = np.zeros(n, dtype=np.uint8)
        count = np.zeros(n, dtype=np.uint32)
        mean = np.zeros((n, 3))
        cov = np.zeros((n, 9))
        check(lib().eskf_map_query(self._h, _d(xyz), C.c_size_t(n),
                                   keys.ctypes.data_as(C.POINTER(C.c_int32)),
                                   hit.ctypes.data_as(C.POINTER(C.c_uint8)),
                                   count.ctypes.data_as(C.POINTER(C.c_uint32)), _d(mean), _d(cov)))
        return keys, hit.astype(bool), count, mean, cov.reshape(n, 3, 3)

    def export(self):
        cap = self.size()
        keys = np.zeros((cap, 3), dtype=np.int32)
        count = np.zeros(cap, dtype=np.uint32)
        mean = np.zeros((cap, 3))
        cov = np.zeros((cap, 9))
        n = C.c_size_t(0)
        check(lib().eskf_map_export(self._h, C.c_size_t(cap), C.byref(n),
                                    keys.ctypes.data_as(C.POINTER(C.c_int32)),
                                    count.ctypes.data_as(C.POINTER(C.c_uint32)), _d(mean), _d(cov)))
        k = n.value
        return keys[:k], count[:k], mean[:k], cov[:k].reshape(k, 3, 3)

    def align_cloud(self, cloud: Cloud, guess, max_iteration=100, translation_sq_threshold=1e-6,
                    cosine_threshold=0.9999, neighbor_mode=1, trace=False):
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        info, bufs = _make_info(max_iteration, trace)
        T = np.zeros(16)
        check(lib().eskf_align_cloud(self.ctx._h, self._h, cloud._h, _d(_f64(guess)), C.byref(prm),
                                     _d(T), C.byref(info)))
        return _info_dict(T, info, bufs)

    def align_cloud_begin(self, cloud: Cloud, guess, max_iteration=100, translation_sq_threshold=1e-6,
                          cosine_threshold=0.9999, neighbor_mode=1):
        """Launch the registration and return; collect it with Context.align_end()."""
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        check(lib().eskf_align_cloud_begin(self.ctx._h, self._h, cloud._h, _d(_f64(guess)), C.byref(prm), None))

    def align_cloud_fixed(self, cloud: Cloud, guess, iterations, neighbor_mode=1, trace=False):
        info, bufs = _make_info(iterations, trace)
        T = np.zeros(16)
        check(lib().eskf_align_cloud_fixed(self.ctx._h, self._h, cloud._h, _d(_f64(guess)),
                                           C.c_int(iterations), C.c_int(neighbor_mode), _d(T),
                                           C.byref(info)))
        return _info_dict(T, info, bufs)

    def align_cloud_p2p(self, cloud: Cloud, guess, comm: "Comm", max_iteration=100,
                        translation_sq_threshold=1e-6, cosine_threshold=0.9999, neighbor_mode=1,
                        fixed_iterations=0, trace=False):
        """COLLECTIVE over comm: this rank's point range against the replicated map; the H/b
        sums meet in NVLink peer mailboxes inside the persistent kernel."""
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        nit = fixed_iterations if fixed_iterations > 0 else max_iteration
        info, bufs = _make_info(nit, trace)
        T = np.zeros(16)
        check(lib().eskf_align_cloud_p2p(self.ctx._h, self._h, cloud._h, _d(_f64(guess)), C.byref(prm),
                                         comm._h, C.c_int(fixed_iterations), _d(T), C.byref(info)))
        return _info_dict(T, info, bufs)

    def p2p_stepper(self, clouds, guess, comm: "Comm", fixed_iterations):
        """step(k): eskf_align_cloud_p2p on clouds[k % len(clouds)] with every argument marshalled once
        (a timed loop then measures the library call, not Python's argument handling).  Returns
        (step, T) where T receives the last pose."""
        prm = IcpParams(100, 1, 1e-6, 0.9999)
        info = AlignInfo()
        T = np.zeros(16)
        g = _f64(guess).copy()
        fn = lib().eskf_align_cloud_p2p
        args = [(self.ctx._h, self._h, c._h, _d(g), C.byref(prm), comm._h, C.c_int(fixed_iterations), _d(T),
                 C.byref(info)) for c in clouds]
        n = len(args)

        def step(k):
            st = fn(*args[k % n])
            if st != OK:
                check(st)

        step._keep = (prm, info, g, clouds)
        return step, T

    def align_cloud_sharded(self, cloud: Cloud, guess, allreduce, max_iteration=100,
                            translation_sq_threshold=1e-6, cosine_threshold=0.9999,
                            neighbor_mode=1, fixed_iterations=0, trace=False):
        """allreduce(device_ptr: int, count: int, stream: int) -> None sums the fp64 buffer
        across ranks in place (e.g. torch.distributed.all_reduce on a tensor view)."""
        prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
        nit = fixed_iterations if fixed_iterations > 0 else max_iteration
        info, bufs = _make_info(nit, trace)
        T = np.zeros(16)
        err = []

        def _cb(_user, buf, count, stream):
            try:
                allreduce(int(buf), int(count), int(stream or 0))
                return 0
            except Exception as e:  # surfaced below
                err.append(e)
                return 1

        cb = ALLREDUCE_FN(_cb) if allreduce is not None else C.cast(None, ALLREDUCE_FN)
        st = lib().eskf_align_cloud_sharded(self.ctx._h, self._h, cloud._h, _d(_f64(guess)),
                                            C.byref(prm), cb, None, C.c_int(fixed_iterations),
                                            _d(T), C.byref(info))
        if err:
            raise err[0]
        check(st)
        return _info_dict(T, info, bufs)


def align_batch(ctxs, maps, clouds, guesses, max_iteration=100, translation_sq_threshold=1e-6,
                cosine_threshold=0.9999, neighbor_mode=1):
    """eskf_align_batch: job i = (maps[i], clouds[i], guesses[i]) runs on ctxs[i % len(ctxs)], to
    which its map and cloud must belong; one registration stays in flight per context.
    Returns one result dict per job."""
    n = len(maps)
    if not (len(clouds) == n and len(guesses) == n):
        raise ValueError("maps, clouds and guesses must have the same length")
    prm = IcpParams(max_iteration, neighbor_mode, translation_sq_threshold, cosine_threshold)
    carr = (C.c_void_p * len(ctxs))(*[c._h.value for c in ctxs])
    marr = (C.c_void_p * max(n, 1))(*[m._h.value for m in maps])
    karr = (C.c_void_p * max(n, 1))(*[c._h.value for c in clouds])
    G = np.ascontiguousarray(np.stack([_f64(g).reshape(16) for g in guesses]) if n else np.zeros((0, 16)))
    T = np.zeros((n, 16))
    infos = (AlignInfo * max(n, 1))()
    check(lib().eskf_align_batch(carr, C.c_int(len(ctxs)), marr, karr, _d(G), C.c_size_t(n), C.byref(prm),
                                 _d(T), infos))
    return [_info_dict(T[i], infos[i], None) for i in range(n)]
