"""C++ oracle vs the independent NumPy/SciPy second oracle
on seeded synthetic data.  Integer / set outputs must agree exactly, floating
point to rounding."""
import numpy as np
import pytest

from eskf_lio_b200 import synth as S
from oracle import np_oracle as NP


@pytest.fixture(scope="module")
def small_scene():
    rng = np.random.default_rng(7)
    scene = S.hall_scene()
    poses = S.arc_trajectory(4)
    scans = [S.make_scan(scene, T, rng) for T in poses]
    return scene, poses, scans


def test_scan_shape_and_time_order(small_scene):
    _, _, scans = small_scene
    xyz, t = scans[0]
    assert xyz.shape == (64000, 3) and t.shape == (64000,)
    assert (np.diff(t) >= 0).all()
    assert (xyz.astype(np.float32).astype(np.float64) == xyz).all()
    r = np.linalg.norm(xyz, axis=1)
    assert r.min() >= 0.25 and r.max() <= 121.0


def test_transform_and_keys_match_numpy(oracle, small_scene):
    _, poses, scans = small_scene
    xyz = scans[0][0][:5000]
    cov = np.repeat(np.diag([1.0, 1.0, 0.01])[None], len(xyz), axis=0)
    T = poses[2] @ S.default_T_il()
    a, ac = oracle.transform_cloud(xyz, cov, T)
    b, bc = NP.transform_cloud(xyz, cov, T)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-13)
    np.testing.assert_allclose(ac, bc, rtol=0, atol=1e-15)
    for v in (0.1, 0.3, 0.5):
        ka, kb = oracle.voxel_index(a, v), NP.voxel_index(a, v)
        np.testing.assert_array_equal(ka, kb)


def test_downsample_cov_matches_scipy(oracle, small_scene):
    _, _, scans = small_scene
    xyz = oracle.transform_cloud(scans[1][0][::4], None, S.default_T_il())[0]
    p1, c1, s1 = oracle.downsample_cov(xyz, 0.5)
    p2, c2, s2 = NP.downsample_cov(xyz, 0.5)
    np.testing.assert_array_equal(s1.astype(np.int64), s2)      # kept set + order bit-exact
    np.testing.assert_array_equal(p1, p2)
    # U F V^T (true SVD) == U F U^T (Jacobi eigen) to rounding / eigen-gap conditioning
    assert np.abs(c1 - c2).max() < 1e-7
    assert np.median(np.abs(c1 - c2)) < 1e-12
    ev = np.linalg.eigvalsh(c1)
    np.testing.assert_allclose(ev, np.tile([1e-2, 1.0, 1.0], (len(ev), 1)), atol=1e-9)


def test_map_and_align_match_numpy(oracle, small_scene):
    _, poses, scans = small_scene
    T_il = S.default_T_il()
    m = oracle.Map(0.5, 1000)
    nm = NP.NpMap(0.5, 1000)
    for (xyz, t), T in zip(scans[:3], poses[:3]):
        p, c, _ = oracle.preprocess(xyz[::2], t[::2], T_il, None, 0.5)
        _, _, pw, cw = m.update(p, c, T, initialize=True)
        nm.insert(*NP.transform_cloud(p, c, T))
    keys, count, mean, cov = m.export()
    nk = sorted(nm.grid)
    assert [tuple(k) for k in keys.tolist()] == nk
    assert count.tolist() == [nm.grid[k][0] for k in nk]
    np.testing.assert_allclose(mean, np.stack([nm.grid[k][1] for k in nk]), atol=1e-12)
    np.testing.assert_allclose(cov, np.stack([nm.grid[k][2] for k in nk]), atol=1e-12)

    p, c, _ = oracle.preprocess(scans[3][0][::2], scans[3][1][::2], T_il, None, 0.5)
    guess = poses[3] @ S.perturbation()
    # one linearisation at the guess: correspondence set exact, H/b to rounding
    pg, cg = oracle.transform_cloud(p, c, guess)
    H, b, hit, nc = m.linearize(pg, cg)
    H2, b2, hit2 = nm.linearize(*NP.transform_cloud(p, c, guess))
    np.testing.assert_array_equal(hit[:, 0], hit2)
    assert nc == hit2.sum() and nc > 1000
    np.testing.assert_allclose(H, H2, rtol=1e-10, atol=1e-6)
    np.testing.assert_allclose(b, b2, rtol=1e-9, atol=1e-7)
    r = m.align(p, c, guess)
    T2, it2, conv2 = nm.align(p, c, guess)
    assert r["iterations"] == it2 and r["converged"] == conv2 and r["converged"]
    np.testing.assert_allclose(r["T"], T2, atol=1e-9)
    err = np.linalg.inv(poses[3]) @ r["T"]
    assert np.linalg.norm(err[:3, 3]) < 0.02  # recovers the known transform to map resolution


def test_direct7_superset_of_direct1(oracle, small_scene):
    _, poses, scans = small_scene
    T_il = S.default_T_il()
    m = oracle.Map(0.5, 1000)
    p, c, _ = oracle.preprocess(scans[0][0][::2], scans[0][1][::2], T_il, None, 0.5)
    m.update(p, c, poses[0], initialize=True)
    q, qc, _ = oracle.preprocess(scans[1][0][::2], scans[1][1][::2], T_il, None, 0.5)
    qw, qcw = oracle.transform_cloud(q, qc, poses[1])
    H1, b1, hit1, n1 = m.linearize(qw, qcw, 1)
    H7, b7, hit7, n7 = m.linearize(qw, qcw, 7)
    np.testing.assert_array_equal(hit7[:, 0], hit1[:, 0])
    assert n7 > n1 and n7 == hit7.sum()
