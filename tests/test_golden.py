"""Committed golden fixtures (tests/golden/hotpath_v1.npz, generated from the
oracle by tests/golden/make_golden.py — the reference has no vectors of its
own): the oracle must keep reproducing them (CPU), and the CUDA path must
match them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _scan(g, k):
    return g[f"raw{k}"].astype(np.float64), g[f"time{k}"]


def test_oracle_reproduces_golden(oracle, gold):
    g = gold
    om = oracle.Map(float(g["voxel"]), 1000)
    for k in range(3):
        xyz, t = _scan(g, k)
        p, c, src = oracle.preprocess(xyz, t, g["T_il"], None, float(g["voxel"]))
        np.testing.assert_array_equal(src, g[f"kept{k}"])
        np.testing.assert_allclose(c, g[f"cov{k}"], atol=1e-12)
        if k < 2:
            om.update(p, c, g["poses"][k], initialize=True)
    keys, count, mean, cov = om.export()
    np.testing.assert_array_equal(keys, g["map_keys"])
    np.testing.assert_array_equal(count, g["map_count"])
    np.testing.assert_allclose(mean, g["map_mean"], atol=1e-12)
    pg, cg = oracle.transform_cloud(p, c, g["guess"])
    np.testing.assert_array_equal(oracle.voxel_index(pg, float(g["voxel"])), g["keys_at_guess"])
    H, b, hit, _ = om.linearize(pg, cg)
    np.testing.assert_array_equal(hit[:, 0], g["lin_hit"])
    np.testing.assert_allclose(H, g["lin_H"], rtol=1e-10)
    r = om.align(p, c, g["guess"])
    assert r["iterations"] == int(g["align_iterations"])
    np.testing.assert_allclose(r["T"], g["align_T"], atol=1e-10)


@pytest.mark.gpu
def test_gpu_matches_golden(gold):
    from eskf_lio_b200 import capi
    g = gold
    v = float(g["voxel"])
    ctx = capi.Context(0)
    gm = capi.Map(ctx, v, 1000, 1 << 12)
    for k in range(3):
        xyz, t = _scan(g, k)
        p, c, src = ctx.preprocess(xyz, t, g["T_il"], None, v)
        np.testing.assert_array_equal(src, g[f"kept{k}"])            # kept set bit-exact
        assert np.abs(c - g[f"cov{k}"]).max() < 1e-7
        if k < 2:
            gm.insert(p, g[f"cov{k}"], g["poses"][k])               # golden covariances: exact map
    keys, count, mean, cov = gm.export()
    np.testing.assert_array_equal(keys, g["map_keys"])              # occupancy bit-exact
    np.testing.assert_array_equal(count.astype(np.uint64), g["map_count"])
    np.testing.assert_array_equal(mean, g["map_mean"])
    np.testing.assert_array_equal(cov, g["map_cov"])
    c = g["cov2"]
    qk, _, _, _, _ = gm.query(np.zeros((1, 3)))
    H, b, hit, nc = ctx.linearize(gm, p, c, T=g["guess"])
    np.testing.assert_array_equal(hit[:, 0], g["lin_hit"])          # correspondence set bit-exact
    assert np.linalg.norm(H - g["lin_H"]) / np.linalg.norm(g["lin_H"]) < 1e-4
    assert np.linalg.norm(b - g["lin_b"]) / np.linalg.norm(g["lin_b"]) < 1e-4
    r = ctx.align(gm, p, c, g["guess"])
    assert r["iterations"] == int(g["align_iterations"])
    np.testing.assert_array_equal(r["ncorr"], g["align_ncorr"])
    E = np.linalg.inv(g["align_T"]) @ r["T"]
    assert np.linalg.norm(E[:3, 3]) < 1e-5
    assert np.arccos(np.clip(0.5 * (np.trace(E[:3, :3]) - 1), -1, 1)) < 1e-5


# ------------------------------------------------- ErrorStateKF + Odometry::run
GOLD_ODOM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "odometry_v1.npz")


def _odom_log(g):
    off = np.concatenate([[0], np.cumsum(g["n"])])
    scans = [(g["xyz"][off[i]:off[i + 1]].astype(np.float64), g["time"][off[i]:off[i + 1]])
             for i in range(len(g["n"]))]
    return scans, g["imu"]


def test_oracle_reproduces_golden_odometry(oracle):
    """tests/golden/odometry_v1.npz (make_golden_odometry.py): filter + call order of Odometry::run."""
    g = dict(np.load(GOLD_ODOM))
    scans, imu = _odom_log(g)
    oracle.set_num_threads(1)
    od = oracle.Odometry(oracle.odom_default_config(map_voxel_size=0.5, preprocess_voxel_size=0.5))
    rec = []
    poses = oracle.run_sequence(od, scans, imu, lambda i, o: rec.append(
        (o.info().last_iterations, o.info().last_inserted, o.info().map_voxels, o.info().last_kept)))
    oracle.set_num_threads(oracle.num_threads())
    np.testing.assert_array_equal(np.array(rec), g["rec"])
    np.testing.assert_allclose(np.stack(poses), g["poses"], atol=1e-10)
    st = od.last_state(with_P=True)
    got = np.concatenate([[st["t"]], st["p"], st["v"], st["q"], st["ba"], st["bg"], st["g"]])
    np.testing.assert_allclose(got, g["state"], atol=1e-9)
    assert np.linalg.norm(st["P"] - g["P"]) / np.linalg.norm(g["P"]) < 1e-9


@pytest.mark.gpu
def test_gpu_odometry_matches_golden():
    """The product's host ESKF + Odometry with the hot path on the B200, against the committed
    oracle trajectory (no liboracle.so involved)."""
    from eskf_lio_b200 import odometry
    from gpu_common import pose_err
    g = dict(np.load(GOLD_ODOM))
    scans, imu = _odom_log(g)
    od = odometry.Odometry(odometry.default_config(map_voxel_size=0.5, preprocess_voxel_size=0.5))
    rec = []
    poses = odometry.run_sequence(od, scans, imu, lambda i, o: rec.append(
        (o.info().last_iterations, o.info().last_inserted, o.info().map_voxels)))
    assert [r[0] for r in rec] == g["rec"][:, 0].tolist()      # Gauss-Newton iteration counts
    assert [r[1] for r in rec] == g["rec"][:, 1].tolist()      # keyframe-gate decisions
    assert [int(r[2]) for r in rec] == g["rec"][:, 2].astype(int).tolist()   # map occupancy after every frame
    for a, b in zip(poses, g["poses"]):
        dt, dr = pose_err(b, a)
        assert dt < 1e-5 and dr < 1e-5
    od.close()
