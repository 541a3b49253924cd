"""Synthetic LiDAR / IMU generator (the Hilti bag the reference is launched on,
launch/eskf_lio.launch.py:11-13, is not available offline).

Sensor model, scenes and configs follow SURVEY.md 8(d): a Hesai-XT32-like
32-beam x 2000-column sweep (64,000 rays, azimuth-major so per-point times
ascend, as CloudPreprocessor.cpp:33,54-60 requires), range noise sigma=0.02 m
so the 30-NN covariances are full rank, points rounded to float32 then widened
(include/ESKF_LIO/Subscriber.hpp:89-95), LiDAR->IMU extrinsics from
config/hilti_config.yaml:20-23.  Host-side NumPy only; the generated arrays are
fed identically to the CPU oracle and the GPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

N_BEAMS = 32
N_COLS = 2000
SWEEP_S = 0.1
RANGE_MIN = 0.3
RANGE_MAX = 120.0
RANGE_SIGMA = 0.02


def quat_xyzw_to_matrix(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def default_T_il():
    """sensors.lidar.extrinsics (config/hilti_config.yaml:20-23), quaternion x,y,z,w."""
    T = np.eye(4)
    T[:3, :3] = quat_xyzw_to_matrix([0.7071068, -0.7071068, 0.0, 0.0])
    T[:3, 3] = [-0.001, -0.00855, 0.055]
    return T


def rotvec_matrix(rv):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-15:
        return np.eye(3)
    a = rv / th
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def pose(xyz, rotvec=(0.0, 0.0, 0.0)):
    T = np.eye(4)
    T[:3, :3] = rotvec_matrix(rotvec)
    T[:3, 3] = xyz
    return T


@dataclass
class Scene:
    """Closed analytic scene: one room (rays always hit its inside), solid
    axis-aligned boxes, and half-space planes (n.x = d, visible from n.x < d)."""
    room_lo: np.ndarray
    room_hi: np.ndarray
    boxes: list = field(default_factory=list)    # (lo[3], hi[3])
    planes: list = field(default_factory=list)   # (n[3] unit, d)

    def raycast(self, origin, dirs):
        """Nearest hit distance and surface normal for every ray."""
        o = np.asarray(origin, dtype=np.float64)
        D = np.asarray(dirs, dtype=np.float64)
        n_rays = D.shape[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / D
            # room: exit distance through the face each component heads to
            bound = np.where(D > 0, self.room_hi, self.room_lo)
            tt = (bound - o) * inv
            tt = np.where(np.isfinite(tt) & (D != 0), tt, np.inf)
            ax = np.argmin(tt, axis=1)
            t_best = tt[np.arange(n_rays), ax]
            normal = np.zeros((n_rays, 3))
            normal[np.arange(n_rays), ax] = -np.sign(D[np.arange(n_rays), ax])
            for lo, hi in self.boxes:
                t1 = (lo - o) * inv
                t2 = (hi - o) * inv
                tn = np.minimum(t1, t2)
                tf = np.maximum(t1, t2)
                tn = np.where(np.isnan(tn), -np.inf, tn)
                tf = np.where(np.isnan(tf), np.inf, tf)
                axn = np.argmax(tn, axis=1)
                t_near = tn[np.arange(n_rays), axn]
                t_far = np.min(tf, axis=1)
                hit = (t_near <= t_far) & (t_near > 1e-9) & (t_near < t_best)
                t_best = np.where(hit, t_near, t_best)
                nb = np.zeros((n_rays, 3))
                nb[np.arange(n_rays), axn] = -np.sign(D[np.arange(n_rays), axn])
                normal = np.where(hit[:, None], nb, normal)
            for n, d in self.planes:
                denom = D @ n
                t = (d - o @ n) / denom
                hit = (denom > 1e-12) & (t > 1e-9) & (t < t_best)
                t_best = np.where(hit, t, t_best)
                normal = np.where(hit[:, None], -n[None, :], normal)
        return t_best, normal

    def sample_surface(self, n, rng):
        """n points uniform (by area) on room faces and box faces, with the
        outward-facing (towards free space) unit normal of each."""
        faces = []  # (origin, u, v, normal)
        lo, hi = self.room_lo, self.room_hi
        for ax in range(3):
            u, v = [a for a in range(3) if a != ax]
            for side, sgn in ((lo, 1.0), (hi, -1.0)):
                org = lo.copy()
                org[ax] = side[ax]
                eu = np.zeros(3)
                eu[u] = hi[u] - lo[u]
                ev = np.zeros(3)
                ev[v] = hi[v] - lo[v]
                nn = np.zeros(3)
                nn[ax] = sgn
                faces.append((org, eu, ev, nn))
        for blo, bhi in self.boxes:
            for ax in range(3):
                u, v = [a for a in range(3) if a != ax]
                for side, sgn in ((blo, -1.0), (bhi, 1.0)):
                    if ax == 2 and sgn < 0 and abs(blo[2] - lo[2]) < 1e-9:
                        continue  # bottom face sits on the floor
                    org = blo.copy()
                    org[ax] = side[ax]
                    eu = np.zeros(3)
                    eu[u] = bhi[u] - blo[u]
                    ev = np.zeros(3)
                    ev[v] = bhi[v] - blo[v]
                    nn = np.zeros(3)
                    nn[ax] = sgn
                    faces.append((org, eu, ev, nn))
        area = np.array([np.linalg.norm(np.cross(f[1], f[2])) for f in faces])
        which = rng.choice(len(faces), size=n, p=area / area.sum())
        a = rng.random(n)
        b = rng.random(n)
        org = np.stack([f[0] for f in faces])[which]
        eu = np.stack([f[1] for f in faces])[which]
        ev = np.stack([f[2] for f in faces])[which]
        nn = np.stack([f[3] for f in faces])[which]
        return org + a[:, None] * eu + b[:, None] * ev, nn


def hall_scene():
    """Config 1/5: hall 60 x 40 x 10 m, 12 box pillars, 2 slanted planes."""
    boxes = []
    for ix, x in enumerate((-22.0, -11.0, 0.5, 11.0, 22.0, -16.0)):
        for y in (-11.0, 9.5):
            w = 0.8 + 0.15 * ((ix + (y > 0)) % 3)
            h = 10.0 if ix % 2 == 0 else 3.0 + 0.7 * ix
            boxes.append((np.array([x - w, y - w, 0.0]), np.array([x + w, y + w, h])))
    n1 = np.array([1.0, 0.35, 0.6])
    n1 /= np.linalg.norm(n1)
    n2 = np.array([-0.5, 1.0, 0.45])
    n2 /= np.linalg.norm(n2)
    planes = [(n1, float(n1 @ np.array([27.0, 15.0, 6.0]))),
              (n2, float(n2 @ np.array([-24.0, 17.0, 7.0])))]
    return Scene(np.array([-30.0, -20.0, 0.0]), np.array([30.0, 20.0, 10.0]), boxes, planes)


def corridor_scene():
    """Config 2: 400 x 30 x 10 m corridor with boxes along both walls (so the
    >100 m eviction of LocalMap.cpp:149-154 triggers on a long run)."""
    boxes = []
    for i in range(40):
        x = -195.0 + 10.0 * i
        side = -1.0 if i % 2 == 0 else 1.0
        d = 1.0 + 0.5 * (i % 4)
        h = 2.5 + 0.6 * (i % 7)
        y0 = side * 15.0
        lo = np.array([x, min(y0, y0 - side * d), 0.0])
        hi = np.array([x + 3.0 + 0.4 * (i % 5), max(y0, y0 - side * d), h])
        boxes.append((lo, hi))
    return Scene(np.array([-200.0, -15.0, 0.0]), np.array([200.0, 15.0, 10.0]), boxes, [])


def block_scene():
    """Configs 3/4: 200 x 200 x 20 m area with a grid of building blocks."""
    boxes = []
    k = 0
    for x in np.arange(-85.0, 86.0, 24.0):
        for y in np.arange(-85.0, 86.0, 24.0):
            if abs(x - 11.0) < 1.0 and abs(y - 11.0) < 1.0:
                continue
            w = 4.0 + (k % 5)
            h = 6.0 + 1.7 * (k % 8)
            boxes.append((np.array([x - w, y - w, 0.0]), np.array([x + w, y + w, h])))
            k += 1
    return Scene(np.array([-100.0, -100.0, 0.0]), np.array([100.0, 100.0, 20.0]), boxes, [])


def beam_dirs():
    """Ray directions in the LiDAR frame, azimuth-major: index = col*32 + beam."""
    el = np.deg2rad(np.arange(-16.0, 16.0, 1.0))                    # 32 beams
    az = np.deg2rad(np.arange(N_COLS) * (360.0 / N_COLS))           # 2000 columns
    ce, se = np.cos(el), np.sin(el)
    d = np.stack([np.outer(np.cos(az), ce), np.outer(np.sin(az), ce),
                  np.broadcast_to(se, (N_COLS, N_BEAMS))], axis=-1)
    return d.reshape(-1, 3)


def make_scan(scene, T_wb, rng, T_il=None, t0=0.0, pose_fn=None, chunk=50):
    """One 64,000-ray sweep.

    T_wb: IMU-body pose in the world used for every ray, unless ``pose_fn(t)``
    is given (motion-distorted sweep: every block of `chunk` columns is cast
    from pose_fn at the block's mid time).
    Returns (xyz_lidar float64[N,3] (float32-rounded), point_time float64[N]).
    """
    T_il = default_T_il() if T_il is None else T_il
    dirs_l = beam_dirs()
    col = np.repeat(np.arange(N_COLS), N_BEAMS)
    times = t0 + SWEEP_S * col / N_COLS
    if pose_fn is None:
        T_wl = T_wb @ T_il
        t_hit, _ = scene.raycast(T_wl[:3, 3], dirs_l @ T_wl[:3, :3].T)
    else:
        t_hit = np.empty(dirs_l.shape[0])
        for c0 in range(0, N_COLS, chunk):
            sl = slice(c0 * N_BEAMS, (c0 + chunk) * N_BEAMS)
            T_wl = pose_fn(t0 + SWEEP_S * (c0 + 0.5 * chunk) / N_COLS) @ T_il
            t_hit[sl], _ = scene.raycast(T_wl[:3, 3], dirs_l[sl] @ T_wl[:3, :3].T)
    r = t_hit + rng.normal(0.0, RANGE_SIGMA, size=t_hit.shape)
    keep = (r >= RANGE_MIN) & (r <= RANGE_MAX) & np.isfinite(r)
    xyz = (dirs_l * r[:, None])[keep]
    xyz = xyz.astype(np.float32).astype(np.float64)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(times[keep])


def arc_trajectory(n, step=0.5, yaw_deg_per_scan=1.0, z=1.5, start=(-8.0, -3.0)):
    """Ground-truth body poses 'step' m apart along a gentle arc (config 1)."""
    poses = []
    x, y, yaw = start[0], start[1], 0.0
    for _ in range(n):
        poses.append(pose([x, y, z], [0.0, 0.0, yaw]))
        yaw += np.deg2rad(yaw_deg_per_scan)
        x += step * np.cos(yaw)
        y += step * np.sin(yaw)
    return poses


def perturbation(dt=(0.10, -0.05, 0.03), angle_deg=1.0, axis=(1.0, 1.0, 1.0)):
    """The config-1 initial-guess perturbation."""
    a = np.asarray(axis, dtype=np.float64)
    a /= np.linalg.norm(a)
    return pose(dt, a * np.deg2rad(angle_deg))


def dense_cloud(scene, n, rng, sigma=RANGE_SIGMA):
    """Config 3: n surface samples + normal noise; covariances I - 0.99 n n^T."""
    p, nn = scene.sample_surface(n, rng)
    p = p + nn * rng.normal(0.0, sigma, size=(n, 1))
    cov = np.eye(3)[None, :, :] - 0.99 * nn[:, :, None] * nn[:, None, :]
    return np.ascontiguousarray(p), np.ascontiguousarray(cov)


# ---------------------------------------------------------------- sequences
# BASELINE.json configs[1]: 10 Hz LiDAR + 400 Hz IMU
# from an analytic C^inf trajectory.  World frame = IMU body frame at t = 0
# (the reference inserts its first scan at Identity, src/Odometry.cpp:61, and
# starts the filter at rest, include/ESKF_LIO/Types.hpp:34-36), so the
# trajectory starts at the identity pose with zero velocity and acceleration.
IMU_RATE = 400.0
GRAVITY_STATE = np.array([0.01165152782783894, -0.008749296634685332, 9.804989173462031])
BIAS_A = np.array([0.06080652138668933, 0.08353074835853214, 0.057072968234636895])
BIAS_G = np.array([-0.0015351229643790084, -0.0013449146576507546, 0.00030127855524786183])


@dataclass
class Trajectory:
    """p(t), R(t) = Rz(yaw) Ry(pitch) Rx(roll), every component A sin(k u(t))
    (x: v u) with the smooth start u(t) = t - tau tanh(t / tau)."""
    v: float = 1.2            # cruise speed along +x [m/s] (0.12 m per frame > the 0.1 m map-update gate)
    tau: float = 1.0
    amp: tuple = (0.0, 1.5, 0.10, 0.25, 0.03, 0.02)     # x(unused) y z yaw pitch roll
    freq: tuple = (0.0, 0.20, 0.50, 0.30, 0.70, 0.90)   # rad per metre-of-u
    T_scene_world: np.ndarray = field(default_factory=lambda: np.eye(4))

    def _u(self, t):
        th = np.tanh(t / self.tau)
        return t - self.tau * th, th * th, 2.0 * th * (1.0 - th * th) / self.tau

    def _comp(self, i, t):
        u, du, ddu = self._u(t)
        if i == 0:
            return self.v * u, self.v * du, self.v * ddu
        A, k = self.amp[i], self.freq[i]
        s, c = np.sin(k * u), np.cos(k * u)
        return A * s, A * k * c * du, -A * k * k * s * du * du + A * k * c * ddu

    def rotation(self, t):
        psi, th, ph = self._comp(3, t)[0], self._comp(4, t)[0], self._comp(5, t)[0]
        cz, sz, cy, sy, cx, sx = np.cos(psi), np.sin(psi), np.cos(th), np.sin(th), np.cos(ph), np.sin(ph)
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
        Ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
        Rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
        return Rz @ Ry @ Rx

    def pose_world(self, t):
        T = np.eye(4)
        T[:3, :3] = self.rotation(t)
        T[:3, 3] = [self._comp(i, t)[0] for i in range(3)]
        return T

    def pose_scene(self, t):
        return self.T_scene_world @ self.pose_world(t)

    def velocity(self, t):
        return np.array([self._comp(i, t)[1] for i in range(3)])

    def imu(self, t, rng=None, gyro_sigma=0.0, accel_sigma=0.0):
        """(gyro, accel) the reference's model would read at time t:
        a_world = R (acc - b_a) + gravity, attitude <- attitude * Exp((w - b_g) dt)
        (src/ErrorStateKF.cpp:86-96)."""
        a_w = np.array([self._comp(i, t)[2] for i in range(3)])
        (psi, dpsi, _), (th, dth, _), (ph, dph, _) = self._comp(3, t), self._comp(4, t), self._comp(5, t)
        w_b = np.array([dph - dpsi * np.sin(th),
                        dth * np.cos(ph) + dpsi * np.sin(ph) * np.cos(th),
                        -dth * np.sin(ph) + dpsi * np.cos(ph) * np.cos(th)])
        f_b = self.rotation(t).T @ (a_w - GRAVITY_STATE)
        gyro, acc = w_b + BIAS_G, f_b + BIAS_A
        if rng is not None:
            gyro = gyro + rng.normal(0.0, gyro_sigma, 3)
            acc = acc + rng.normal(0.0, accel_sigma, 3)
        return gyro, acc


def corridor_trajectory():
    """Config 2: start 50 m from the corridor's -x end wall, 1.5 m above the floor."""
    return Trajectory(T_scene_world=pose([-150.0, 0.0, 1.5]))


def hall_trajectory():
    """Short sequences (tests, default bench): the config-1 hall."""
    return Trajectory(amp=(0.0, 1.0, 0.10, 0.25, 0.03, 0.02), T_scene_world=pose([-20.0, -2.0, 1.5]))


def make_sequence(scene, traj, n_frames, seed, imu_noise=True, chunk=50):
    """n_frames motion-distorted sweeps + the IMU stream covering them.

    Returns (scans, imu) with scans[i] = (xyz_lidar[N,3], point_time[N]) for the
    sweep starting at 0.1 i s, and imu = float64[M, 7] rows (t, gyro xyz, accel xyz)
    at 400 Hz from t = 0 to 0.1 n_frames + 0.05 s.
    """
    rng = np.random.default_rng(seed)
    T_il = default_T_il()
    scans = []
    for i in range(n_frames):
        scans.append(make_scan(scene, None, rng, T_il, t0=SWEEP_S * i, pose_fn=traj.pose_scene, chunk=chunk))
    n_imu = int(round((SWEEP_S * n_frames + 0.05) * IMU_RATE)) + 1
    imu = np.empty((n_imu, 7))
    # noise densities of config/hilti_config.yaml:13-16 at 400 Hz
    gs = 0.014 * np.sqrt(IMU_RATE) * np.pi / 180.0 if imu_noise else 0.0
    as_ = 120e-6 * 9.81 * np.sqrt(IMU_RATE) if imu_noise else 0.0
    for k in range(n_imu):
        t = k / IMU_RATE
        g, a = traj.imu(t, rng if imu_noise else None, gs, as_)
        imu[k, 0] = t
        imu[k, 1:4] = g
        imu[k, 4:7] = a
    return scans, imu
