"""One checker, two subjects: an implementation of the hot path (the CPU oracle on the CPU
suite, the CUDA path through the C ABI on the GPU suite) held DIRECTLY against the reference's
own sources compiled on the API shims (oracle/ref.py, `seq` build).  The oracle adapter keeps the
checker itself tested where there is no GPU."""
import numpy as np

from eskf_lio_b200 import synth as S
from gpu_common import pose_err


class OracleImpl:
    def __init__(self, oracle):
        self.O = oracle

    def preprocess(self, xyz, t, T_il, states, voxel):
        return self.O.preprocess(xyz, t, T_il, states, voxel)[:2]

    def new_map(self, voxel):
        m = self.O.Map(voxel, 1000)
        m.set_update_params(remove_enabled=False)
        return m

    def insert(self, m, p, c, T):
        m.update(p, c, T, initialize=True)

    def export(self, m):
        k, n, mean, cov = m.export()
        return k, n.astype(np.uint64), mean, cov

    def hits(self, m, xyz):
        return m.query(xyz)[1]

    def align(self, m, p, c, guess):
        r = m.align(p, c, guess)
        return r["T"], bool(r["converged"])


class GpuImpl:
    def __init__(self, capi, ctx):
        self.capi, self.ctx = capi, ctx

    def preprocess(self, xyz, t, T_il, states, voxel):
        return self.ctx.preprocess(xyz, t, T_il, states, voxel)[:2]

    def new_map(self, voxel):
        return self.capi.Map(self.ctx, voxel, 1000, 1 << 14)

    def insert(self, m, p, c, T):
        m.insert(p, c, T)

    def export(self, m):
        k, n, mean, cov = m.export()
        return k, n.astype(np.uint64), mean, cov

    def hits(self, m, xyz):
        return m.query(xyz)[1]

    def align(self, m, p, c, guess):
        r = self.ctx.align(m, p, c, guess)
        return r["T"], bool(r["converged"])


def rows_sorted(p, c):
    o = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    return p[o], c[o]


def check_against_reference(ref, impl, frames, sweep, cov_tol, pose_tol):
    """ref: oracle.ref.Ref('seq'); frames: gpu_common.Frames; sweep: (xyz, t, states)."""
    # CloudPreprocessor::process with deskew: the kept set, point for point
    xyz, t, states = sweep
    T_il = S.default_T_il()
    for st in (None, states):
        p, c = impl.preprocess(xyz, t, T_il, st, 0.5)
        rp, rc = ref.preprocess(xyz, t, T_il, st, 0.5)
        assert len(p) == len(rp) > 1000
        (p, c), (rp, rc) = rows_sorted(p, c), rows_sorted(rp, rc)
        np.testing.assert_array_equal(p, rp)
        assert np.abs(c - rc).max() < cov_tol
    # LocalMap::updateLocalMap: keys, counts, running means and covariances bit for bit
    m, rm = impl.new_map(0.5), ref.Map(voxel_map=0.5, remove_enabled=0)
    for (p, c), T in zip(frames.ds[:4], frames.poses[:4]):
        impl.insert(m, p, c, T)
        rm.update(p, c, T, initialize=True)
    for a, b in zip(impl.export(m), rm.export()):
        np.testing.assert_array_equal(a, b)
    # LocalMap::correspondenceMatching at the perturbed guess: the correspondence set
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    wp, wc = ref.transform_cloud(p, c, guess)
    sp, _, _, _ = rm.correspondences(wp, wc)
    np.testing.assert_array_equal(wp[impl.hits(m, wp)], sp)
    # ICP::align: the final pose and the converged flag
    T, conv = impl.align(m, p, c, guess)
    Tr, rconv = rm.align(p, c, guess)
    assert conv == rconv is True
    dt, dr = pose_err(Tr, T)
    assert dt < pose_tol and dr < pose_tol, (dt, dr)
