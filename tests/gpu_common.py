"""Shared seeded data for the GPU parity tests (oracle = checker only)."""
import numpy as np

from eskf_lio_b200 import synth as S


class Frames:
    """A few synthetic scans preprocessed by the ORACLE (so that map / align
    tests do not depend on the GPU preprocessor being right)."""

    def __init__(self, oracle, n_scans=5, seed=11, decim=2, voxel=0.5):
        rng = np.random.default_rng(seed)
        self.scene = S.hall_scene()
        self.poses = S.arc_trajectory(n_scans)
        self.T_il = S.default_T_il()
        self.voxel = voxel
        self.raw = []
        self.ds = []
        for T in self.poses:
            xyz, t = S.make_scan(self.scene, T, rng)
            xyz, t = xyz[::decim].copy(), t[::decim].copy()
            self.raw.append((xyz, t))
            p, c, src = oracle.preprocess(xyz, t, self.T_il, None, voxel)
            self.ds.append((p, c))

    def build_maps(self, oracle, capi, ctx, n, voxel=None, cap=1000, hint=1 << 14):
        voxel = voxel or self.voxel
        om = oracle.Map(voxel, cap)
        gm = capi.Map(ctx, voxel, cap, hint)
        for (p, c), T in zip(self.ds[:n], self.poses[:n]):
            om.update(p, c, T, initialize=True)
            gm.insert(p, c, T)
        return om, gm


def rel_err(a, b):
    """norm-wise relative error (H/b parity is norm-wise)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def pose_err(A, B):
    E = np.linalg.inv(A) @ B
    R = E[:3, :3]
    # atan2 form: arccos loses half the digits near the identity
    sin = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.linalg.norm(E[:3, 3])), float(np.arctan2(sin, 0.5 * (np.trace(R) - 1.0)))
