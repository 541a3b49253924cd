"""Independent NumPy/SciPy SECOND ORACLE (test infrastructure only).

Written against the reference's source text with library routines
(np.linalg.inv/solve/svd, scipy cKDTree, scipy Rotation) instead of the
hand-restated arithmetic in eskf_oracle.cpp, so the two can pin each other:
they agree to rounding (1e-9 or tighter) and exactly on every integer/set
output.  PARITY UNPINNED by the reference itself (it has no tests).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation


def skew(v):
    # src/Utils.cpp:5-11
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def compute_J(r):
    # src/Utils.cpp:40-54
    r = np.asarray(r, dtype=np.float64)
    ang = np.linalg.norm(r)
    if ang < 1e-6:
        return np.eye(3)
    a = r / ang
    f1 = np.sin(ang) / ang
    f2 = (1.0 - np.cos(ang)) / ang
    return f1 * np.eye(3) + (1.0 - f1) * np.outer(a, a) + f2 * skew(a)


def se3_to_SE3(se3):
    # src/Utils.cpp:56-63
    se3 = np.asarray(se3, dtype=np.float64)
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec(se3[3:]).as_matrix()
    T[:3, 3] = compute_J(se3[3:]) @ se3[:3]
    return T


def voxel_index(xyz, v):
    # src/LocalMap.cpp:114-118
    return np.floor(np.asarray(xyz, dtype=np.float64) / v).astype(np.int32)


def transform_cloud(xyz, cov, T):
    # Open3D PointCloud::Transform
    R, t = T[:3, :3], T[:3, 3]
    xyz2 = xyz @ R.T + t
    cov2 = None if cov is None else np.einsum("ij,njk,lk->nil", R, cov, R)
    return xyz2, cov2


def jtj_jtr(p, mu, C):
    # src/Registration.cpp:83-102
    J = np.hstack([np.eye(3), -skew(p)])
    W = np.linalg.inv(C)
    JT = J.T @ W
    return JT @ J, JT @ (p - mu)


class NpMap:
    """dict-based LocalMap (src/LocalMap.cpp:47-58, LocalMap.hpp:72-87)."""

    def __init__(self, voxel_size, cap=1000):
        self.v = voxel_size
        self.cap = cap
        self.grid = {}

    def insert(self, xyz, cov):
        keys = voxel_index(xyz, self.v)
        for p, c, k in zip(xyz, cov, keys):
            k = (int(k[0]), int(k[1]), int(k[2]))
            e = self.grid.get(k)
            if e is None:
                self.grid[k] = [1, p.copy(), c.copy()]
            elif e[0] < self.cap:
                n = e[0]
                e[1] = (n * e[1] + p) / (n + 1)
                e[2] = (n * e[2] + c) / (n + 1)
                e[0] = n + 1

    def evict(self, pos, thr):
        # src/LocalMap.cpp:149-154
        dead = [k for k in self.grid
                if np.linalg.norm((np.array(k, dtype=np.float64) + 0.5) * self.v - pos) > thr]
        for k in dead:
            del self.grid[k]
        return len(dead)

    def linearize(self, xyz, cov):
        # src/LocalMap.cpp:78-112 + src/Registration.cpp:56-76
        H = np.zeros((6, 6))
        b = np.zeros(6)
        hit = np.zeros(len(xyz), dtype=bool)
        keys = voxel_index(xyz, self.v)
        for i, (p, c, k) in enumerate(zip(xyz, cov, keys)):
            e = self.grid.get((int(k[0]), int(k[1]), int(k[2])))
            if e is None:
                continue
            hit[i] = True
            Hi, bi = jtj_jtr(p, e[1], c + e[2])
            H += Hi
            b += bi
        return H, b, hit

    def align(self, xyz, cov, guess, max_iteration=100, trans_sq_thr=1e-6, cos_thr=0.9999):
        # src/Registration.cpp:7-35
        total = guess.copy()
        pts, covs = transform_cloud(xyz, cov, guess)
        it = 0
        converged = False
        for _ in range(max_iteration):
            H, b, _hit = self.linearize(pts, covs)
            if np.linalg.matrix_rank(H) == 6:
                se3 = np.linalg.solve(H, -b)
            else:
                se3 = np.linalg.lstsq(H, -b, rcond=None)[0]
            step = se3_to_SE3(se3)
            total = step @ total
            it += 1
            cosine = 0.5 * (np.trace(step[:3, :3]) - 1.0)
            if cosine >= cos_thr and step[:3, 3] @ step[:3, 3] <= trans_sq_thr:
                converged = True
                break
            pts, covs = transform_cloud(pts, covs, step)
        return total, it, converged


def downsample_cov(xyz, voxel_size, k=30):
    """src/CloudPreprocessor.cpp:76-127 with cKDTree + a true SVD.
    Output in ascending source index order."""
    xyz = np.asarray(xyz, dtype=np.float64)
    keys = voxel_index(xyz, voxel_size).astype(np.int64)
    first = {}
    for i, kk in enumerate(map(tuple, keys)):
        first.setdefault(kk, i)
    src = np.array(sorted(first.values()), dtype=np.int64)
    tree = cKDTree(xyz)
    kk = min(k, len(xyz))
    _, nn = tree.query(xyz[src], k=kk)
    nn = nn.reshape(len(src), kk)
    F = np.diag([1.0, 1.0, 1e-2])
    covs = np.zeros((len(src), 3, 3))
    for j in range(len(src)):
        if kk >= 3:
            P = xyz[nn[j]]
            m = P.mean(axis=0)
            Cm = (P.T @ P) / kk - np.outer(m, m)
        else:
            Cm = np.eye(3)
        U, _s, Vt = np.linalg.svd(Cm)
        covs[j] = U @ F @ Vt
    return xyz[src], covs, src


def slerp(qa, qb, t):
    """Eigen Quaterniond::slerp, quaternions as (x,y,z,w)."""
    d = float(np.dot(qa, qb))
    ad = abs(d)
    if ad >= 1.0 - np.finfo(np.float64).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        s0 = np.sin((1.0 - t) * th) / np.sin(th)
        s1 = np.sin(t * th) / np.sin(th)
    if d < 0:
        s1 = -s1
    return s0 * qa + s1 * qb


def deskew(xyz, times, ts, pos, quat):
    """src/CloudPreprocessor.cpp:25-74, vectorised per segment, including the
    'last segment left untouched' behaviour."""
    xyz = np.array(xyz, dtype=np.float64)
    ts = np.asarray(ts, dtype=np.float64)
    end_time = times[-1]
    before = np.nonzero(ts <= end_time)[0][-1]
    after = min(before + 1, len(ts) - 1)
    f = (end_time - ts[before]) / (ts[after] - ts[before] + 1e-6)
    q = slerp(np.asarray(quat[before]), np.asarray(quat[after]), f)
    Rend = _quat_matrix(q)  # Eigen toRotationMatrix does not renormalise
    tend = pos[before] + f * (pos[after] - pos[before])
    Rinv = Rend.T
    tinv = -Rinv @ tend
    start = 0
    for s in range(after + 1):
        later = np.nonzero(times[start:] >= ts[s])[0]
        if len(later) == 0:
            break  # reference: pointEndIndex never advances again
        end = start + later[0]
        if end == start:
            continue
        Rs = _quat_matrix(np.asarray(quat[s]))
        R = Rinv @ Rs
        t = Rinv @ pos[s] + tinv
        xyz[start:end] = xyz[start:end] @ R.T + t
        start = end
    return xyz


def _quat_matrix(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
