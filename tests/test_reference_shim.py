"""The ORACLE against the REFERENCE'S OWN SOURCES.

/root/reference/src/{Registration,LocalMap,CloudPreprocessor,Utils,ErrorStateKF,Odometry}.cpp are compiled
where they lie (oracle/Makefile `ref`) against from-scratch shims of the Eigen / Open3D / yaml-cpp
API subset they use (oracle/refshim/include; the real libraries do not exist here) and driven
through oracle/refshim/ref_capi.cpp.  This pins the oracle's restatement of the reference's
control flow and formulas — loop structure, thresholds, gates, the deskew quirk, first-point-per-
voxel, addPoint order and cap, J^T W J / J^T W r, the LDLT step, convergence, the filter's
predict / update / reset — to the reference's text; the third-party arithmetic underneath is
the shims' (restated from the libraries' published algorithms, like the oracle's).

Two builds of the same sources: `seq` sums short inner products left to right (the oracle's
documented convention) => results must be BIT-IDENTICAL wherever no parallel summation order is
involved; `tree` pairs them as Eigen's unrolled reductions do (a0 + (a1 + a2)) => results agree to
a few ulp and every discrete outcome (voxel keys, kept sets, correspondence sets, gates,
convergence, iteration counts) is unchanged.

Skipped where neither /root/reference nor the prebuilt oracle/_ref libraries exist."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

import oracle as O
from oracle import ref as R
from eskf_lio_b200 import synth as S
from gpu_common import Frames, pose_err

pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources / prebuilt oracle/_ref not present")

KINDS = ["seq", "tree"]


@pytest.fixture(scope="module", params=KINDS)
def ref(request):
    return R.Ref(request.param)


@pytest.fixture(scope="module")
def frames():
    return Frames(O, n_scans=5, seed=11, decim=4, voxel=0.5)


def exact(ref):
    return ref.kind == "seq"


def close(a, b, ref, ulps=64):
    """bit-identical in the `seq` build, within `ulps` units of the largest magnitude in `tree`"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if exact(ref):
        np.testing.assert_array_equal(a, b)
    else:
        scale = max(float(np.abs(b).max()), 1e-300)
        assert float(np.abs(a - b).max()) <= ulps * np.finfo(np.float64).eps * scale


# ------------------------------------------------------------------ Utils.cpp
def test_utils_match(ref):
    rng = np.random.default_rng(3)
    for _ in range(300):
        v = rng.normal(size=3)
        np.testing.assert_array_equal(ref.skew(v), O.skew(v))
        rv = rng.normal(size=3) * rng.choice([0.0, 1e-9, 1e-7, 1e-3, 0.5, 3.0])
        close(ref.rotvec_to_matrix(rv), O.rotvec_to_matrix(rv), ref)
        se3 = rng.normal(size=6) * rng.choice([0.0, 1e-9, 1e-7, 1e-3, 0.5])   # both branches of computeJ
        close(ref.se3_to_SE3(se3), O.se3_to_SE3(se3), ref)
        Rm = S.rotvec_matrix(rng.normal(size=3) * rng.choice([1e-9, 1e-6, 0.1, 3.0, 3.14159]))
        a, b = ref.rotation_matrix_to_vector(Rm), O.rotation_matrix_to_vector(Rm)
        if exact(ref):
            np.testing.assert_array_equal(a, b)
        else:
            assert np.abs(a - b).max() < 1e-12


def test_interpolate_se3_matches(ref):
    rng = np.random.default_rng(4)
    for _ in range(100):
        t1 = rng.uniform(0, 10)
        t2 = t1 + rng.uniform(1e-4, 0.1)
        p1, p2 = rng.normal(size=3), rng.normal(size=3)
        q1 = Rot.from_rotvec(rng.normal(size=3)).as_quat()
        q2 = (Rot.from_quat(q1) * Rot.from_rotvec(rng.normal(size=3) * rng.choice([1e-9, 0.01, 1.0]))).as_quat()
        if rng.random() < 0.3:
            q2 = -q2                                     # slerp's d < 0 branch
        t = rng.uniform(t1, t2)
        close(ref.interpolate_SE3((t1, p1, q1), (t2, p2, q2), t),
              O.interpolate_SE3((t1, p1, q1), (t2, p2, q2), t), ref)


# ---------------------------------------- Open3D Transform, getVoxelIndex
def test_transform_and_voxel_index_match(ref):
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(2000, 3)) * 30.0
    A = rng.normal(size=(2000, 3, 3))
    cov = A @ A.transpose(0, 2, 1)
    T = np.eye(4)
    T[:3, :3] = S.rotvec_matrix([0.1, -0.2, 0.3])
    T[:3, 3] = [1.0, -2.0, 3.0]
    rp, rc = ref.transform_cloud(xyz, cov, T)
    op, oc = O.transform_cloud(xyz, cov, T)
    np.testing.assert_array_equal(rp, op)                 # (4x4 homogeneous product: same order in both builds)
    close(rc, oc, ref)
    edge = np.array([[-0.1, -0.5, -0.0], [0.5, 0.4999999999999999, 0.3], [0.6, 0.9, -0.3], [1e-300, -1e-300, 0.0]])
    for v in (0.1, 0.3, 0.5, 1.0):
        np.testing.assert_array_equal(ref.voxel_index(xyz, v), O.voxel_index(xyz, v))
        np.testing.assert_array_equal(ref.voxel_index(edge, v), O.voxel_index(edge, v))


# ------------------------------------------------------------ LocalMap.cpp
def build_maps(ref, frames, n, cap=1000, **over):
    om = O.Map(0.5, cap)
    om.set_update_params(remove_enabled=False)
    rm = ref.Map(voxel_map=0.5, max_points_per_voxel=cap, remove_enabled=0, **over)
    for (p, c), T in zip(frames.ds[:n], frames.poses[:n]):
        _, _, ox, oc = om.update(p, c, T, initialize=True)
        rx, rc = rm.update(p, c, T, initialize=True)
        np.testing.assert_array_equal(rx, ox)             # the caller's cloud ends up in the world frame
        close(rc, oc, ref)
    return om, rm


@pytest.mark.parametrize("cap", [1000, 3])
def test_map_update_matches(ref, frames, cap):
    om, rm = build_maps(ref, frames, 4, cap)
    assert rm.size() == om.size()
    ok, oc, omean, ocov = om.export()
    rk, rc, rmean, rcov = rm.export()
    np.testing.assert_array_equal(rk, ok)
    np.testing.assert_array_equal(rc, oc)                  # numPoints, incl. the hard cap
    np.testing.assert_array_equal(rmean, omean)
    close(rcov, ocov, ref)
    assert int(oc.max()) == (cap if cap == 3 else int(oc.max()))


def test_keyframe_gate_matches(ref, frames):
    om, rm = build_maps(ref, frames, 1)
    rng = np.random.default_rng(6)
    prev = frames.poses[0]
    for _ in range(200):
        # around both thresholds (cos 0.985 <-> 9.94 deg, |t|^2 = 1e-2)
        d = S.perturbation(dt=tuple(rng.normal(size=3) * 0.06), angle_deg=float(rng.uniform(0, 20)))
        cur = prev @ d
        assert rm.needs_map_update(prev, cur) == O.needs_map_update(prev, cur, 1e-2, 0.985)
    # gated-out frames must leave the map untouched but still move prevTransform_ (LocalMap.cpp:39-42)
    p, c = frames.ds[1]
    small = prev @ S.perturbation(dt=(0.01, 0.0, 0.0), angle_deg=0.1)
    n0 = rm.size()
    ins, _, _, _ = om.update(p, c, small, initialize=False)
    rm.update(p, c, small, initialize=False)
    assert not ins and rm.size() == n0 == om.size()


def test_eviction_matches(ref, frames):
    om = O.Map(0.5, 1000)
    om.set_update_params(remove_enabled=False)
    # removing_period 0: the reference sweeps on every inserting call (its clock is omp_get_wtime())
    rm = ref.Map(voxel_map=0.5, remove_enabled=1, remove_distance=12.0, remove_period=0.0)
    for (p, c), T in zip(frames.ds[:3], frames.poses[:3]):
        om.update(p, c, T, initialize=True)
        om.evict(T[:3, 3], 12.0)
        rm.update(p, c, T, initialize=True)
        assert rm.size() == om.size()
    np.testing.assert_array_equal(rm.export()[0], om.export()[0])


def test_correspondences_match(ref, frames):
    om, rm = build_maps(ref, frames, 4)
    p, c = frames.ds[4]
    wp, wc = O.transform_cloud(p, c, frames.poses[4] @ S.perturbation())
    _, hit, _, mean, cov = om.query(wp)
    sp, sc, mp, mc = rm.correspondences(wp, wc)
    assert len(sp) == int(hit.sum()) > 500
    np.testing.assert_array_equal(sp, wp[hit])             # the correspondence SET, point for point
    np.testing.assert_array_equal(sc, wc[hit])
    np.testing.assert_array_equal(mp, mean[hit])
    close(mc, cov[hit], ref)


# -------------------------------------------------------- Registration.cpp
def test_jtj_jtr_matches(ref, frames):
    om, rm = build_maps(ref, frames, 4)
    p, c = frames.ds[4]
    wp, wc = O.transform_cloud(p, c, frames.poses[4] @ S.perturbation())
    _, hit, _, mean, cov = om.query(wp)
    for i in np.nonzero(hit)[0][:300]:
        Ho, bo = O.jtj_jtr(wp[i], mean[i], wc[i] + cov[i])
        Hr, br = rm.jtj_jtr(wp[i], mean[i], wc[i] + cov[i])
        close(Hr, Ho, ref, ulps=256)
        close(br, bo, ref, ulps=4096)                      # (b is a difference of nearly equal terms)


def test_gauss_newton_step_and_align_match(ref, frames):
    om, rm = build_maps(ref, frames, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    ro = om.align(p, c, guess)
    assert ro["converged"] and 2 <= ro["iterations"] <= 20
    # every iteration: the reference's correspondenceMatching + computeTransform on the oracle's
    # working cloud gives the oracle's step (sums differ only in OpenMP chunking: 1e-16 relative)
    wp, wc = O.transform_cloud(p, c, guess)
    for k in range(ro["iterations"]):
        Ts, nc = rm.gn_step(wp, wc)
        assert nc == int(ro["ncorr"][k])
        assert np.abs(Ts - ro["step"][k]).max() < 1e-12
        assert rm.convergence_check(Ts) == (k == ro["iterations"] - 1)
        wp, wc = O.transform_cloud(wp, wc, ro["step"][k])
    Tr, conv = rm.align(p, c, guess)
    assert conv
    dt, dr = pose_err(ro["T"], Tr)
    assert dt < 1e-12 and dr < 1e-12
    # no correspondences at all: LDLT of a zero matrix solves to zero => identity step, "converged"
    far = p + np.array([1e4, 0.0, 0.0])
    To, Tr = om.align(far, c, np.eye(4)), rm.align(far, c, np.eye(4))
    np.testing.assert_array_equal(Tr[0], To["T"])
    assert Tr[1] and To["converged"] and To["iterations"] == 1
    # max_iteration exhausted: last estimate, not converged
    o1 = om.align(p, c, guess, max_iteration=2)
    r1 = ref.Map(cfg=R.default_config(voxel_map=0.5, remove_enabled=0, max_iteration=2))
    for (pp, cc), T in zip(frames.ds[:4], frames.poses[:4]):
        r1.update(pp, cc, T, initialize=True)
    T1, c1 = r1.align(p, c, guess)
    assert not c1 and not o1["converged"] and o1["iterations"] == 2
    assert max(pose_err(o1["T"], T1)) < 1e-12


# ---------------------------------------------------- CloudPreprocessor.cpp
def sweep_and_states(seed, late_states):
    rng = np.random.default_rng(seed)
    scene, poses = S.hall_scene(), S.arc_trajectory(3)
    xyz, t = S.make_scan(scene, poses[1], rng)
    xyz, t = xyz[::4].copy(), t[::4].copy()
    # 400 Hz filter states, `late_states` of them after the sweep's last point (the reference needs
    # at least one: CloudPreprocessor.cpp:45).  For the first state after the sweep the scan for "the
    # first point not older than the state" runs off the end, so the tail of the sweep stays
    # untransformed (CloudPreprocessor.cpp:54-65): the quirk the oracle has to reproduce
    n = int(np.ceil((t[-1] - t[0] + 0.01) / 0.0025)) + late_states
    ts = t[0] - 0.01 + np.arange(n) * 0.0025
    pos = np.stack([0.5 * (ts - ts[0]), 0.1 * np.sin(3 * ts), 0.02 * ts], 1)
    quat = Rot.from_rotvec(np.stack([0.02 * np.sin(2 * ts), 0.01 * ts, 0.3 * (ts - ts[0])], 1)).as_quat()
    return xyz, t, (ts, pos, quat)


def rows_sorted(p, c):
    o = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    return p[o], c[o]


@pytest.mark.parametrize("case", ["no_states", "three_states_after_sweep", "one_state_after_sweep"])
@pytest.mark.parametrize("voxel", [0.5, 0.3])
def test_preprocess_matches(ref, case, voxel):
    xyz, t, states = sweep_and_states(7, late_states=3 if case == "three_states_after_sweep" else 1)
    if case == "no_states":
        states = None
    T_il = S.default_T_il()
    op, oc, _ = O.preprocess(xyz, t, T_il, states, voxel)
    rp, rc = ref.preprocess(xyz, t, T_il, states, voxel)
    assert len(rp) == len(op) > 1000                        # the kept SET (first point per voxel)
    (op, oc), (rp, rc) = rows_sorted(op, oc), rows_sorted(rp, rc)
    if exact(ref) or states is None:
        np.testing.assert_array_equal(rp, op)
    else:
        assert np.abs(rp - op).max() < 1e-12                 # Isometry * point: inner order of three
    # 30-NN covariance + U diag(1,1,0.01) V^T (two different Jacobi solvers: ~1e-9 where two
    # singular values nearly coincide, 1e-15 typically)
    assert np.abs(rc - oc).max() < 5e-8
    # (in the `tree` build the deskewed points themselves differ by ~1e-14, which the regularisation amplifies)
    med = 1e-13 if (exact(ref) or states is None) else 1e-11
    assert np.median(np.abs(rc - oc).reshape(len(rc), -1).max(axis=1)) < med


def unsorted_stamps(t, seed=3):
    """Ring-major style stamps: the sweep's stamps permuted in blocks, so they are not monotonic
    (the reference's deskew, src/CloudPreprocessor.cpp:54-61, scans them linearly whatever their order)."""
    rng = np.random.default_rng(seed)
    blocks = np.array_split(np.arange(len(t)), 37)
    order = np.concatenate([blocks[k] for k in rng.permutation(len(blocks))])
    return t[order].copy()


def test_preprocess_with_unsorted_stamps_matches(ref):
    """per-point stamps that are not non-decreasing.  The oracle (and the CUDA path, see
    tests/test_gpu_parity.py) must follow the reference's forward linear scan, not a binary search."""
    xyz, t, states = sweep_and_states(7, late_states=3)
    tu = unsorted_stamps(t)
    assert np.any(np.diff(tu) < 0)
    T_il = S.default_T_il()
    op, oc, _ = O.preprocess(xyz, tu, T_il, states, 0.5)
    rp, rc = ref.preprocess(xyz, tu, T_il, states, 0.5)
    assert len(rp) == len(op) > 1000
    (op, oc), (rp, rc) = rows_sorted(op, oc), rows_sorted(rp, rc)
    if exact(ref):
        np.testing.assert_array_equal(rp, op)
    else:
        assert np.abs(rp - op).max() < 1e-12
    # and the result differs from the sorted-stamp one: the test exercises a different segmentation
    sp, _, _ = O.preprocess(xyz, t, T_il, states, 0.5)
    assert len(sp) != len(op) or np.abs(rows_sorted(sp, np.zeros((len(sp), 3, 3)))[0] - op).max() > 1e-6


def test_range_crop_definition(ref):
    """The range crop the oracle defines (the reference has none): equal to running the reference's
    transform + deskew on the whole sweep, erasing the out-of-range points, then the reference's own
    downsample + covariance step on what is left; off by default."""
    xyz, t, states = sweep_and_states(7, late_states=3)
    T_il = S.default_T_il()
    r = np.linalg.norm(xyz, axis=1)
    lo, hi = float(np.quantile(r, 0.2)), float(np.quantile(r, 0.8))
    op, oc, osrc = O.preprocess(xyz, t, T_il, states, 0.5, min_range=lo, max_range=hi)
    assert len(op) > 500 and np.all(r[osrc] >= lo) and np.all(r[osrc] <= hi)
    p0, _, s0 = O.preprocess(xyz, t, T_il, states, 0.5)
    assert not np.all((r[s0] >= lo) & (r[s0] <= hi))            # the uncropped run keeps out-of-range points
    # composition on the reference side: its own deskew of the full sweep (voxel size tiny: every point
    # survives the reference's downsample step, giving the deskewed positions), then its downsample on the rest
    keep = ((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2] >= lo * lo) & \
           ((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2] <= hi * hi)
    pd = O.deskew(O.transform_cloud(xyz, None, T_il)[0], t, states)
    rp, rc = ref.preprocess(pd[keep], t[keep], np.eye(4), None, 0.5)
    assert len(rp) == len(op)
    (a, ac), (b, bc) = rows_sorted(op, oc), rows_sorted(rp, rc)
    np.testing.assert_array_equal(a, b)
    assert np.abs(ac - bc).max() < 5e-8
    # max_range 0 = unbounded; (0, 0) = off
    q, _, qs = O.preprocess(xyz, t, T_il, states, 0.5, min_range=lo)
    assert np.all(r[qs] >= lo) and np.any(r[qs] > hi)
    z, _, zs = O.preprocess(xyz, t, T_il, states, 0.5, min_range=0.0, max_range=0.0)
    np.testing.assert_array_equal(zs, s0)


# ----------------------------------------------------------- ErrorStateKF.cpp
def test_filter_predict_update_match(ref, frames):
    om, rm = build_maps(ref, frames, 4)
    rng = np.random.default_rng(8)
    # IMU at 400 Hz for 0.1 s, consistent with standing still at the pose of frame 4 ... roughly:
    # the point is identical inputs, not realism
    g = np.array(R.default_config().gravity)
    ts = 0.0025 * np.arange(1, 45)
    gyro = 0.05 * rng.normal(size=(len(ts), 3)) + np.array(R.default_config().bias_g)
    acc = -g + 0.2 * rng.normal(size=(len(ts), 3)) + np.array(R.default_config().bias_a)
    lidar_end = float(ts[39]) + 1e-4                          # 4 samples lie beyond it: rolled back

    kf = ref.Eskf()
    kf.initialize(0.0)
    od, twin = O.Odometry(), O.Odometry()
    kf.process(-1.0, gyro[0], acc[0])                         # dt < 0: ignored (ErrorStateKF.cpp:77-79)
    od.kf_process(-1.0, gyro[0], acc[0])
    for i, t in enumerate(ts):
        kf.process(float(t), gyro[i], acc[i])
        od.kf_process(float(t), gyro[i], acc[i])
        if t <= lidar_end:
            twin.kf_process(float(t), gyro[i], acc[i])
    assert kf.num_states() == len(ts) + 1 == od.info().n_states

    def same_state(tol_P):
        a, b = kf.state(-1, with_P=True), od.last_state(with_P=True)
        assert a["timestamp"] == b["t"]
        for x, y in (("position", "p"), ("velocity", "v"), ("attitude_xyzw", "q"), ("bias_a", "ba"),
                     ("bias_g", "bg"), ("gravity", "g")):
            assert np.abs(a[x] - b[y]).max() < 1e-12, x
        assert np.abs(a["P"] - b["P"]).max() <= tol_P * np.abs(b["P"]).max()

    same_state(1e-12)                                         # 44 predict steps: F P F^T + F_i Q F_i^T

    # update(): rollback, ICP from the predicted pose, Kalman update, injection, reset
    p, c = frames.ds[4]
    s = twin.last_state()
    guess = np.eye(4)
    guess[:3, :3] = O.quat_to_matrix(s["q"])
    guess[:3, 3] = s["p"]
    # (the filter starts at the origin while the map was built along the arc: register the cloud of
    #  frame 4 moved into the filter's frame so that ICP has something to converge to)
    p4, c4 = O.transform_cloud(p, c, np.linalg.inv(guess) @ frames.poses[4] @ S.perturbation())
    obs = om.align(p4, c4, guess)
    assert obs["converged"]
    T_ref = kf.update(rm, p4, c4, lidar_end)
    g_or, T_or = od.kf_update_with_observation(lidar_end, obs["T"])
    np.testing.assert_allclose(g_or, guess, rtol=0, atol=0)
    assert kf.num_states() == 41 + 1 == od.info().n_states     # 4 states rolled back, 1 appended
    assert max(pose_err(T_or, T_ref)) < 1e-11
    same_state(1e-9)


# ------------------------------------ the committed golden fixture (what the GPU test is held to)
def test_reference_sources_reproduce_golden_fixture(ref):
    """tests/golden/hotpath_v1.npz was generated from the oracle; the reference's own sources (on the
    shims) reproduce it: kept sets, covariances, map, correspondence set at the guess, final pose."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.npz")))
    v = float(g["voxel"])
    rm = ref.Map(voxel_map=v, remove_enabled=0)
    for k in range(3):
        xyz, t = g[f"raw{k}"].astype(np.float64), g[f"time{k}"]
        rp, rc = ref.preprocess(xyz, t, g["T_il"], None, v)
        # golden rows are in ascending source index: the kept points are the T_il-transformed raw ones
        gp = O.transform_cloud(xyz, None, g["T_il"])[0][g[f"kept{k}"]]
        assert len(rp) == len(gp)
        (rp, rc), (gp_s, gc_s) = rows_sorted(rp, rc), rows_sorted(gp, g[f"cov{k}"])
        np.testing.assert_array_equal(rp, gp_s)                       # kept set, bit for bit
        assert np.abs(rc - gc_s).max() < 1e-7
        if k < 2:
            rm.update(gp, g[f"cov{k}"], g["poses"][k], initialize=True)   # golden covariances: exact map
    keys, count, mean, cov = rm.export()
    np.testing.assert_array_equal(keys, g["map_keys"])
    np.testing.assert_array_equal(count, g["map_count"])
    np.testing.assert_array_equal(mean, g["map_mean"])
    close(cov, g["map_cov"], ref)
    pg, cg = O.transform_cloud(gp, g["cov2"], g["guess"])
    np.testing.assert_array_equal(ref.voxel_index(pg, v), g["keys_at_guess"])
    sp, _, _, _ = rm.correspondences(pg, cg)
    np.testing.assert_array_equal(sp, pg[g["lin_hit"].astype(bool)])  # correspondence set
    T, conv = rm.align(gp, g["cov2"], g["guess"])
    # (this sparse fixture exhausts max_iteration = 100: "ICP not converged!", last estimate returned)
    assert int(g["align_iterations"]) == 100 and not conv
    assert max(pose_err(g["align_T"], T)) < 1e-10


# ----------------------------------------------------------------- Odometry.cpp
def test_reference_odometry_run_reproduces_golden_trajectory(ref):
    """ESKF_LIO::Odometry::run itself (src/Odometry.cpp, with the reference's ErrorStateKF,
    CloudPreprocessor, ICP and LocalMap) over the committed 20-frame log: the trajectory, the final
    filter state and covariance and the map occupancy of tests/golden/odometry_v1.npz — which the
    oracle generated and the GPU path is held to — come out of the reference's own loop."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "odometry_v1.npz")))
    off = np.concatenate([[0], np.cumsum(g["n"])])
    scans = [(g["xyz"][off[i]:off[i + 1]].astype(np.float64), g["time"][off[i]:off[i + 1]])
             for i in range(len(g["n"]))]
    poses, state, voxels, n_states = ref.run_odometry(scans, g["imu"], float(g["state"][0]),
                                                      voxel_map=0.5, voxel_pre=0.5)
    assert len(poses) == len(g["poses"]) == 20
    for a, b in zip(g["poses"], poses):
        assert max(pose_err(a, b)) < 1e-9
    assert voxels == int(g["rec"][-1, 2])                              # map occupancy after the last frame
    got = np.concatenate([[state["timestamp"]], state["position"], state["velocity"], state["attitude_xyzw"],
                          state["bias_a"], state["bias_g"], state["gravity"]])
    np.testing.assert_allclose(got, g["state"], atol=1e-9)
    assert np.linalg.norm(state["P"] - g["P"]) / np.linalg.norm(g["P"]) < 1e-9


# ------------------------------------------- the checker the GPU suite runs, on the oracle
def test_pipeline_checker_on_the_oracle(frames):
    """tests/ref_check.py holds an implementation directly against the reference's sources; the
    GPU suite runs it on the CUDA path (test_gpu_parity.py), here it runs on the oracle."""
    from ref_check import OracleImpl, check_against_reference
    check_against_reference(R.Ref("seq"), OracleImpl(O), frames, sweep_and_states(7, 2), cov_tol=1e-9, pose_tol=1e-12)


def test_full_size_sweep_matches(ref):
    """BASELINE.json's frame: a full 32 x 2000 sweep, deskewed against 400 Hz states, 0.3 m voxels."""
    rng = np.random.default_rng(21)
    xyz, t = S.make_scan(S.hall_scene(), S.arc_trajectory(3)[1], rng)
    assert len(xyz) == 64000
    n = int(np.ceil((t[-1] - t[0] + 0.01) / 0.0025)) + 2
    ts = t[0] - 0.01 + np.arange(n) * 0.0025
    pos = np.stack([0.5 * (ts - ts[0]), 0.1 * np.sin(3 * ts), 0.02 * ts], 1)
    quat = Rot.from_rotvec(np.stack([0.02 * np.sin(2 * ts), 0.01 * ts, 0.3 * (ts - ts[0])], 1)).as_quat()
    T_il = S.default_T_il()
    op, oc, _ = O.preprocess(xyz, t, T_il, (ts, pos, quat), 0.3)
    rp, rc = ref.preprocess(xyz, t, T_il, (ts, pos, quat), 0.3)
    assert len(op) == len(rp) > 10000
    (op, oc), (rp, rc) = rows_sorted(op, oc), rows_sorted(rp, rc)
    if exact(ref):
        np.testing.assert_array_equal(rp, op)
        assert np.abs(rc - oc).max() < 1e-12
    else:
        assert np.abs(rp - op).max() < 1e-12
        assert np.abs(rc - oc).max() < 5e-8


def test_gpu_adapter_of_the_checker_on_a_stand_in(frames):
    """The GPU suite's adapter (ref_check.GpuImpl) driven by a stand-in with capi's call shapes and
    dtypes built on the oracle: keeps the adapter's plumbing tested where there is no GPU."""
    from ref_check import GpuImpl, check_against_reference

    class FakeMap:
        def __init__(self, ctx, voxel, cap, hint):
            self.m = O.Map(voxel, cap)
            self.m.set_update_params(remove_enabled=False)

        def insert(self, p, c, T):
            self.m.update(p, c, T, initialize=True)

        def export(self):
            k, n, mean, cov = self.m.export()
            return k, n.astype(np.uint32), mean, cov

        def query(self, xyz):
            k, hit, n, mean, cov = self.m.query(xyz)
            return k, hit, n.astype(np.uint32), mean, cov

    class FakeCtx:
        def preprocess(self, xyz, t, T_il, states, voxel):
            return O.preprocess(xyz, t, T_il, states, voxel)

        def align(self, m, p, c, guess):
            return m.m.align(p, c, guess)

    class FakeCapi:
        Map = FakeMap

    check_against_reference(R.Ref("seq"), GpuImpl(FakeCapi, FakeCtx()), frames, sweep_and_states(7, 2),
                            cov_tol=1e-6, pose_tol=1e-5)


def test_default_configs_equal_the_references_yaml():
    """config/hilti_config.yaml, read where it lies: the drop-in's defaults (host Config via
    eskf_odom_default_config), the oracle's and the shim wrapper's are the reference's."""
    import os
    import yaml
    path = os.path.join(R.REFERENCE_ROOT, "config", "hilti_config.yaml")
    if not os.path.exists(path):
        pytest.skip("/root/reference is not present on this box")
    y = yaml.safe_load(open(path))
    imu = y["sensors"]["imu"]["intrinsics"]["parameters"]
    lid = y["sensors"]["lidar"]["extrinsics"]
    lm, reg = y["local_map"], y["registration"]
    want = {
        "imu_update_rate": y["sensors"]["imu"]["update_rate"], "bias_a": imu["bias_a"], "bias_g": imu["bias_g"],
        "gravity": imu["gravity"], "accel_noise_density": imu["accel_noise_density"],
        "accel_zero_g_offset": imu["accel_zero_g_offset"], "gyro_noise_density": imu["gyro_noise_density"],
        "gyro_zero_rate_offset": imu["gyro_zero_rate_offset"],
        "translation_noise": y["kalman_filter"]["update"]["translation_noise"],
        "rotation_noise": y["kalman_filter"]["update"]["rotation_noise"],
        "lidar_quaternion_xyzw": lid["quaternion"], "lidar_translation": lid["translation"],
        "map_voxel_size": lm["voxel_size"], "max_points_per_voxel": lm["max_num_points_per_voxel"],
        "update_translation_sq_threshold": lm["update"]["translation_sq_threshold"],
        "update_cosine_threshold": lm["update"]["cosine_threshold"],
        "remove_enabled": int(lm["remove_distant_points"]["enabled"]),
        "remove_distance_threshold": lm["remove_distant_points"]["distance_threshold"],
        "remove_period": lm["remove_distant_points"]["removing_period"],
        "preprocess_voxel_size": y["cloud_preprocessor"]["voxel_size"], "max_iteration": reg["max_iteration"],
        "icp_translation_sq_threshold": reg["translation_sq_threshold"], "icp_cosine_threshold": reg["cosine_threshold"],
    }

    def same(cfg, names):
        for k, v in want.items():
            got = getattr(cfg, names.get(k, k))
            got = list(got) if hasattr(got, "__len__") else got
            assert got == (list(map(float, v)) if isinstance(v, list) else v), k

    from eskf_lio_b200 import odometry
    same(odometry.default_config(), {})                                   # the product's host classes
    same(O.odom_default_config(), {})                                     # the oracle
    same(R.default_config(), {                                            # the wrapper over the reference's classes
        "imu_update_rate": "imu_rate", "lidar_quaternion_xyzw": "lidar_quat_xyzw", "lidar_translation": "lidar_trans",
        "map_voxel_size": "voxel_map", "update_translation_sq_threshold": "update_tsq",
        "update_cosine_threshold": "update_cos", "remove_distance_threshold": "remove_distance",
        "preprocess_voxel_size": "voxel_pre", "icp_translation_sq_threshold": "icp_tsq",
        "icp_cosine_threshold": "icp_cos"})
