"""Known-answer tests pinning the CPU oracle.

The reference has no tests of its own (parity unpinned), so these analytic
cases are the pins.  Reference lines cited per test."""
import numpy as np
import pytest


def test_voxel_index_floor_semantics(oracle):
    # LocalMap.cpp:114-118: floor(p / v), true division, truncating cast
    pts = np.array([[-0.1, -0.5, -0.0], [0.5, 0.4999999999999999, 0.0], [1.0, -1.0, 0.3]])
    assert oracle.voxel_index(pts, 0.5).tolist() == [[-1, -1, 0], [1, 0, 0], [2, -2, 0]]
    assert oracle.voxel_index(pts, 0.1).tolist() == [
        [-1, -5, 0], [5, 4, 0], [10, -10, int(np.floor(0.3 / 0.1))]]
    # 0.3/0.1 is 2.9999999999999996 in fp64: true division must give 2, not 3
    assert oracle.voxel_index([[0.3, 0.6, 0.9]], 0.1).tolist() == [
        [int(np.floor(0.3 / 0.1)), int(np.floor(0.6 / 0.1)), int(np.floor(0.9 / 0.1))]]
    assert oracle.voxel_index([[0.3, 0.6, 0.9]], 0.3).tolist() == [[1, 2, 3]]


def test_add_point_running_mean_and_cap(oracle):
    # LocalMap.hpp:72-87
    m = oracle.Map(1.0, 3)
    pts = np.array([[0.1, 0.2, 0.3], [0.3, 0.2, 0.1], [0.5, 0.8, 0.4], [0.9, 0.9, 0.9]])
    covs = np.stack([np.eye(3) * (i + 1) for i in range(4)])
    m.insert(pts, covs)
    keys, count, mean, cov = m.export()
    assert keys.tolist() == [[0, 0, 0]] and count.tolist() == [3]  # 4th point dropped by the cap
    np.testing.assert_allclose(mean[0], pts[:3].mean(axis=0), rtol=1e-15)
    np.testing.assert_allclose(cov[0], np.eye(3) * 2.0, rtol=1e-15)
    # exact op order: ((n*mean)+p)/(n+1)
    m1 = (1 * pts[0] + pts[1]) / 2
    m2 = (2 * m1 + pts[2]) / 3
    assert (mean[0] == m2).all()


def test_se3_to_SE3_known_answers(oracle):
    # Utils.cpp:56-63
    np.testing.assert_array_equal(oracle.se3_to_SE3(np.zeros(6)), np.eye(4))
    T = oracle.se3_to_SE3([1.0, -2.0, 3.0, 0, 0, 0])
    np.testing.assert_array_equal(T[:3, 3], [1.0, -2.0, 3.0])
    np.testing.assert_array_equal(T[:3, :3], np.eye(3))
    T = oracle.se3_to_SE3([0, 0, 0, 0, 0, np.pi / 2])
    np.testing.assert_allclose(T[:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)
    # theta < 1e-6 branch: J = I exactly
    T = oracle.se3_to_SE3([0.5, 0.25, -1.0, 3e-7, 0, 0])
    np.testing.assert_array_equal(T[:3, 3], [0.5, 0.25, -1.0])
    np.testing.assert_array_equal(oracle.compute_J([0, 0, 9.9e-7]), np.eye(3))
    assert not np.array_equal(oracle.compute_J([0, 0, 1.1e-6]), np.eye(3))


def test_jtj_jtr_closed_forms(oracle):
    # Registration.cpp:83-102 with C = I: H = J^T J ; p == mu: b = 0
    p = np.array([1.5, -2.0, 0.7])
    H, b = oracle.jtj_jtr(p, p, np.eye(3))
    S = oracle.skew(p)
    J = np.hstack([np.eye(3), -S])
    np.testing.assert_allclose(H, J.T @ J, atol=1e-14)
    np.testing.assert_array_equal(b, np.zeros(6))
    np.testing.assert_allclose(H[:3, :3], np.eye(3), atol=0)
    np.testing.assert_allclose(H[3:, 3:], S.T @ S, atol=1e-14)


def test_b_is_gradient_of_cost(oracle):
    """b = J^T W r is the gradient of 0.5 r^T W r under a left perturbation
    T <- exp(xi) T, state ordering [rho ; phi] (Registration.cpp:92-100)."""
    rng = np.random.default_rng(0)
    p = rng.normal(size=3) * 3
    mu = p + rng.normal(size=3) * 0.1
    A = rng.normal(size=(3, 3))
    Cm = A @ A.T + np.eye(3)
    _, b = oracle.jtj_jtr(p, mu, Cm)
    W = np.linalg.inv(Cm)

    def cost(xi):
        T = oracle.se3_to_SE3(xi)
        r = T[:3, :3] @ p + T[:3, 3] - mu
        return 0.5 * r @ W @ r

    g = np.zeros(6)
    h = 1e-6
    for i in range(6):
        e = np.zeros(6)
        e[i] = h
        g[i] = (cost(e) - cost(-e)) / (2 * h)
    np.testing.assert_allclose(b, g, rtol=1e-6, atol=1e-8)


def test_convergence_check_thresholds(oracle):
    # Registration.cpp:37-50: cosine >= thr AND |t|^2 <= thr
    T = np.eye(4)
    assert oracle.convergence_check(T, 1e-6, 0.9999)
    T[:3, 3] = [1e-3, 0, 0]  # |t|^2 == 1e-6 exactly representable product? compare both sides
    assert oracle.convergence_check(T, T[0, 3] ** 2, 0.9999)          # equality passes (not >)
    assert not oracle.convergence_check(T, np.nextafter(T[0, 3] ** 2, 0), 0.9999)
    R = oracle.rotvec_to_matrix([0, 0, 0.02])
    T = np.eye(4)
    T[:3, :3] = R
    c = 0.5 * (np.trace(R) - 1.0)
    assert oracle.convergence_check(T, 1e-6, c)                       # equality passes (not <)
    assert not oracle.convergence_check(T, 1e-6, np.nextafter(c, 1.0))


def test_ldlt_solve(oracle):
    rng = np.random.default_rng(1)
    A = rng.normal(size=(6, 6))
    H = A @ A.T + 0.1 * np.eye(6)
    b = rng.normal(size=6)
    np.testing.assert_allclose(oracle.ldlt_solve6(H, b), np.linalg.solve(H, b), rtol=1e-10)
    # zero system -> zero step (Eigen LDLT pseudo-solve)
    np.testing.assert_array_equal(oracle.ldlt_solve6(np.zeros((6, 6)), b), np.zeros(6))
    # rank-deficient PSD: minimum-norm-like behaviour on the range, no NaN
    v = rng.normal(size=(6, 2))
    Hs = v @ v.T
    x = oracle.ldlt_solve6(Hs, Hs @ np.ones(6))
    assert np.isfinite(x).all()
    np.testing.assert_allclose(Hs @ x, Hs @ np.ones(6), atol=1e-9)


def test_align_identity_and_zero_correspondence(oracle):
    rng = np.random.default_rng(2)
    pts = rng.uniform(-5, 5, size=(500, 3))
    covs = np.repeat(np.eye(3)[None], 500, axis=0)
    m = oracle.Map(0.5, 1000)
    m.insert(pts, covs)
    # map built from the same cloud with <=1 point per voxel -> r = 0 for those; 1 iteration, identity
    r = m.align(pts, covs, np.eye(4))
    assert r["iterations"] == 1 and r["converged"]
    np.testing.assert_allclose(r["T"], np.eye(4), atol=1e-9)
    # zero correspondences: returns the guess, "converged" after one zero step
    g = np.eye(4)
    g[:3, 3] = [1000.0, 1000.0, 1000.0]
    r = m.align(pts, covs, g)
    assert r["iterations"] == 1 and r["converged"] and r["ncorr"].tolist() == [0]
    np.testing.assert_array_equal(r["T"], g)


def test_eviction_strict_greater(oracle):
    # LocalMap.cpp:149-154: erase iff distance > threshold
    m = oracle.Map(1.0, 10)
    pts = np.array([[0.5, 0.5, 0.5], [3.5, 0.5, 0.5], [4.5, 0.5, 0.5]])
    m.insert(pts, np.repeat(np.eye(3)[None], 3, axis=0))
    assert m.evict([0.5, 0.5, 0.5], 3.0) == 1  # centres at distance 0, 3 (kept: not >), 4 (erased)
    keys, _, _, _ = m.export()
    assert keys.tolist() == [[0, 0, 0], [3, 0, 0]]


def test_update_gating_and_first_eviction(oracle):
    # LocalMap.cpp:39-42,60,74 and LocalMap.hpp:40
    m = oracle.Map(0.5, 1000)
    m.set_update_params(1e-2, 0.985, True, 100.0, 10.0)
    pts = np.array([[1.0, 2.0, 0.5], [150.0, 0.0, 0.0]])
    covs = np.repeat(np.eye(3)[None], 2, axis=0)
    ins, removed, w, _ = m.update(pts, covs, np.eye(4), initialize=True, now=0.0)
    assert ins and removed == 1  # currentRemoveTime_ starts at lowest(): first insert always evicts
    T = np.eye(4)
    T[0, 3] = 0.05  # |t|^2 = 2.5e-3 < 1e-2, no rotation -> gated out
    ins, removed, w, _ = m.update(pts, covs, T, now=1.0)
    assert not ins and m.size() == 1
    np.testing.assert_allclose(w[0], [1.05, 2.0, 0.5])  # cloud is transformed even when gated out
    T[0, 3] = 0.12  # vs PREVIOUS FRAME (0.05): 0.07^2 < 1e-2 -> still gated out
    ins, _, _, _ = m.update(pts, covs, T, now=2.0)
    assert not ins
    T[0, 3] = 0.30  # 0.18^2 > 1e-2 -> inserted; 2 s < period -> no eviction
    ins, removed, _, _ = m.update(pts, covs, T, now=3.0)
    assert ins and removed == 0 and m.size() == 2
    keys, count, _, _ = m.export()
    assert keys.tolist() == [[2, 4, 1], [300, 0, 0]] and count.tolist() == [2, 1]


def test_knn_kdtree_matches_bruteforce(oracle):
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(3000, 3)) * [5, 5, 0.5]
    q = pts[::37]
    i1, d1 = oracle.knn(pts, q, 30)
    i2, d2 = oracle.knn(pts, q, 30, bruteforce=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    # fewer points than k: padded with -1
    i3, _ = oracle.knn(pts[:5], pts[:2], 30)
    assert (i3[:, 5:] == -1).all() and sorted(i3[0, :5].tolist()) == [0, 1, 2, 3, 4]


def test_knn_tie_order_on_a_lattice(oracle):
    """Equidistant neighbours: KD-tree and brute force agree on the (distance, index) order."""
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3) * 0.125
    pts = g[np.random.default_rng(4).permutation(len(g))].astype(np.float64)
    i1, d1 = oracle.knn(pts, pts[:150], 30)
    i2, d2 = oracle.knn(pts, pts[:150], 30, bruteforce=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    # ties resolved by ascending index
    for row_i, row_d in zip(i1[:20], d1[:20]):
        for a in range(29):
            if row_d[a] == row_d[a + 1]:
                assert row_i[a] < row_i[a + 1]


def test_regularize_cov(oracle):
    # CloudPreprocessor.cpp:120-123 == I - 0.99 n n^T for a PSD input
    rng = np.random.default_rng(4)
    A = rng.normal(size=(3, 3))
    Cm = A @ A.T
    w, V = np.linalg.eigh(Cm)
    n = V[:, 0]
    np.testing.assert_allclose(oracle.regularize_cov(Cm), np.eye(3) - 0.99 * np.outer(n, n),
                               atol=1e-12)
    np.testing.assert_allclose(oracle.regularize_cov(np.eye(3)), np.diag([1, 1, 1e-2]), atol=1e-15)


def test_deskew_identity_motion_is_noop_and_last_segment_quirk(oracle):
    from oracle import np_oracle as NP
    rng = np.random.default_rng(5)
    n = 4000
    pts = rng.normal(size=(n, 3)) * 10
    times = np.linspace(10.0, 10.1, n)
    ts = 10.0 - 0.005 + 0.0025 * np.arange(46)  # 400 Hz states around the sweep
    pos = np.stack([0.5 * (ts - 10.0), 0.1 * (ts - 10.0) ** 2, np.zeros_like(ts)], axis=1)
    quat = np.stack([np.zeros_like(ts), np.zeros_like(ts), np.sin(0.1 * (ts - 10)),
                     np.cos(0.1 * (ts - 10))], axis=1)
    out = oracle.deskew(pts, times, (ts, pos, quat))
    ref = NP.deskew(pts, times, ts, pos, quat)
    np.testing.assert_allclose(out, ref, atol=1e-12)
    # points after the last state <= end time are left untouched (CloudPreprocessor.cpp:54-65)
    last_state = ts[ts <= times[-1]][-1]
    tail = times >= last_state
    assert tail.sum() > 0
    np.testing.assert_array_equal(out[tail], pts[tail])
    assert not np.array_equal(out[~tail], pts[~tail])
    # zero motion: every segment transform is identity up to rounding
    pos0 = np.zeros_like(pos)
    quat0 = np.tile([0.0, 0.0, 0.0, 1.0], (len(ts), 1))
    np.testing.assert_allclose(oracle.deskew(pts, times, (ts, pos0, quat0)), pts, atol=1e-12)
    with pytest.raises(RuntimeError):
        oracle.deskew(pts, times, (ts + 100.0, pos, quat))
