"""Pins the oracle's ErrorStateKF / Odometry restatement (oracle/odom_oracle.cpp;
reference: src/ErrorStateKF.cpp, src/Odometry.cpp) with an independent NumPy /
SciPy restatement of the same equations and analytic known answers.  CPU only."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from eskf_lio_b200 import synth as S

G_MAG = 9.81


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


class NumpyESKF:
    """src/ErrorStateKF.cpp written with numpy / scipy (Rotation for every
    rotation conversion) — shares no code with the oracle."""

    def __init__(self, cfg):
        self.t = 0.0
        self.p = np.zeros(3)
        self.v = np.zeros(3)
        self.R = Rotation.identity()
        self.ba = np.array(cfg.bias_a)
        self.bg = np.array(cfg.bias_g)
        self.g = np.array(cfg.gravity)
        self.P = 1e-3 * np.eye(18)
        sr = np.sqrt(cfg.imu_update_rate)
        sa = np.array(cfg.accel_noise_density) * G_MAG * sr                  # :29-32
        sg = cfg.gyro_noise_density * sr * np.pi / 180.0                      # :33
        saw = cfg.accel_zero_g_offset * sr * 1e-3 * G_MAG                     # :34
        sgw = cfg.gyro_zero_rate_offset * sr * np.pi / 180.0                  # :35
        self.Q = np.diag(np.concatenate([sa ** 2, [sg ** 2] * 3, [saw ** 2] * 3, [sgw ** 2] * 3]))
        self.Fi = np.zeros((18, 12))
        self.Fi[3:15, :] = np.eye(12)                                         # :45-46
        self.V = np.diag([cfg.translation_noise] * 3 + [cfg.rotation_noise] * 3)
        self.H = np.zeros((6, 18))
        self.H[0:3, 0:3] = np.eye(3)
        self.H[3:6, 6:9] = np.eye(3)

    def process(self, t, gyro, acc):                                          # :76-113
        dt = t - self.t
        if dt < 0:
            return
        R = self.R.as_matrix()
        a = np.asarray(acc) - self.ba
        w = np.asarray(gyro) - self.bg
        dR = Rotation.from_rotvec(w * dt)
        aw = R @ a + self.g
        p = self.p + self.v * dt + 0.5 * aw * dt * dt
        v = self.v + aw * dt
        Qi = self.Q.copy()
        Qi[0:6, 0:6] *= dt * dt
        Qi[6:12, 6:12] *= dt
        F = np.eye(18)
        F[0:3, 3:6] = np.eye(3) * dt
        F[3:6, 6:9] = -R @ skew(a) * dt
        F[3:6, 9:12] = -R * dt
        F[3:6, 15:18] = np.eye(3) * dt
        F[6:9, 6:9] = dR.inv().as_matrix()
        F[6:9, 12:15] = -np.eye(3) * dt
        self.P = F @ self.P @ F.T + self.Fi @ Qi @ self.Fi.T
        self.p, self.v, self.R, self.t = p, v, self.R * dR, t

    def update(self, t, obs):                                                 # :115-162
        guess_R, guess_t = self.R.as_matrix(), self.p.copy()
        res = np.concatenate([obs[:3, 3] - guess_t,
                              Rotation.from_matrix(guess_R.T @ obs[:3, :3]).as_rotvec()])
        K = self.P @ self.H.T @ np.linalg.inv(self.H @ self.P @ self.H.T + self.V)
        e = K @ res
        P = (np.eye(18) - K @ self.H) @ self.P
        self.p = self.p + e[0:3]
        self.v = self.v + e[3:6]
        self.R = self.R * Rotation.from_rotvec(e[6:9])
        self.ba = self.ba + e[9:12]
        self.bg = self.bg + e[12:15]
        self.g = self.g + e[15:18]
        G = np.eye(18)
        G[6:9, 6:9] = np.eye(3) - 0.5 * skew(e[6:9])
        self.P = G @ P @ G.T
        self.t = t


def _state_close(st, ref, tol):
    np.testing.assert_allclose(st["p"], ref.p, rtol=0, atol=tol)
    np.testing.assert_allclose(st["v"], ref.v, rtol=0, atol=tol)
    q = ref.R.as_quat()
    if np.dot(q, st["q"]) < 0:
        q = -q
    np.testing.assert_allclose(st["q"], q, rtol=0, atol=tol)
    np.testing.assert_allclose(st["ba"], ref.ba, rtol=0, atol=tol)
    np.testing.assert_allclose(st["bg"], ref.bg, rtol=0, atol=tol)
    np.testing.assert_allclose(st["g"], ref.g, rtol=0, atol=tol)


def test_process_matches_numpy(oracle):
    """400 IMU steps of ErrorStateKF::process: state and P against the NumPy restatement."""
    cfg = oracle.odom_default_config()
    od = oracle.Odometry(cfg)
    ref = NumpyESKF(cfg)
    tr = S.hall_trajectory()
    rng = np.random.default_rng(3)
    for k in range(1, 401):
        t = k / 400.0
        g, a = tr.imu(t, rng, 0.005, 0.02)
        od.kf_process(t, g, a)
        ref.process(t, g, a)
    st = od.last_state(with_P=True)
    _state_close(st, ref, 1e-9)
    assert np.linalg.norm(st["P"] - ref.P) / np.linalg.norm(ref.P) < 1e-10
    assert np.allclose(st["P"], st["P"].T, rtol=1e-9, atol=1e-12)


def test_update_matches_numpy(oracle):
    """ErrorStateKF::update with a supplied ICP pose: gain, injection, reset, P."""
    cfg = oracle.odom_default_config()
    od = oracle.Odometry(cfg)
    ref = NumpyESKF(cfg)
    tr = S.hall_trajectory()
    t = 0.0
    for frame in range(3):
        for _ in range(40):
            t += 1 / 400.0
            g, a = tr.imu(t)
            od.kf_process(t, g, a)
            ref.process(t, g, a)
        obs = tr.pose_world(t) @ S.perturbation(dt=(0.004, -0.003, 0.002), angle_deg=0.05)
        guess_ref = np.eye(4)
        guess_ref[:3, :3] = ref.R.as_matrix()
        guess_ref[:3, 3] = ref.p
        guess, T = od.kf_update_with_observation(t, obs)
        ref.update(t, obs)
        np.testing.assert_allclose(guess, guess_ref, rtol=0, atol=1e-9)
        st = od.last_state(with_P=True)
        _state_close(st, ref, 1e-8)
        assert np.linalg.norm(st["P"] - ref.P) / np.linalg.norm(ref.P) < 1e-8
        # observation noise 1e-6 against a huge predicted covariance: the filter lands on the ICP pose
        assert np.linalg.norm(T[:3, 3] - obs[:3, 3]) < 1e-4


def test_rotation_matrix_to_vector_kats(oracle):
    """Utils::rotationMatrixToVector (src/Utils.cpp:22-26) incl. both Shepperd branches."""
    assert np.array_equal(oracle.rotation_matrix_to_vector(np.eye(3)), np.zeros(3))
    rng = np.random.default_rng(0)
    for ang in (1e-9, 1e-4, 0.3, 1.5, 2.5, 3.1):   # trace > 0 and trace <= 0
        for _ in range(4):
            ax = rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            rv = oracle.rotation_matrix_to_vector(Rotation.from_rotvec(ang * ax).as_matrix())
            np.testing.assert_allclose(rv, ang * ax, rtol=0, atol=1e-9 + 1e-7 * (ang > 3))
    R = Rotation.from_rotvec([0, 0, np.pi / 2]).as_matrix()
    np.testing.assert_allclose(oracle.rotation_matrix_to_vector(R), [0, 0, np.pi / 2], atol=1e-15)


def test_negative_dt_is_ignored(oracle):
    od = oracle.Odometry()
    od.kf_process(0.01, [0, 0, 0.1], [0, 0, -9.8])
    n = od.info().n_states
    od.kf_process(0.005, [0, 0, 0.1], [0, 0, -9.8])   # :80-82 dt < 0 -> return
    assert od.info().n_states == n


def test_imu_model_integrates_to_trajectory():
    """The synthetic IMU is consistent with the reference's propagation model."""
    tr = S.hall_trajectory()
    dt = 1 / 400.0
    p, v, R = np.zeros(3), np.zeros(3), np.eye(3)
    for k in range(2000):
        g, a = tr.imu(k * dt)
        aw = R @ (a - S.BIAS_A) + S.GRAVITY_STATE
        p = p + v * dt + 0.5 * aw * dt * dt
        v = v + aw * dt
        R = R @ S.rotvec_matrix((g - S.BIAS_G) * dt)
    T = tr.pose_world(2000 * dt)
    assert np.linalg.norm(p - T[:3, 3]) < 5e-3
    assert np.linalg.norm(R - T[:3, :3]) < 1e-3
    assert np.allclose(tr.pose_world(0.0), np.eye(4)) and np.allclose(tr.velocity(0.0), 0.0)


@pytest.fixture(scope="module")
def short_sequence():
    tr = S.hall_trajectory()
    scans, imu = S.make_sequence(S.hall_scene(), tr, 22, seed=5)
    scans = [(x[::2].copy(), t[::2].copy()) for x, t in scans]
    return tr, scans, imu


def test_odometry_tracks_ground_truth(oracle, short_sequence):
    """Odometry::run call order on the oracle: init frame, IMU-predicted guesses, deskew against the
    filter states, keyframe gate; the estimate stays within a few cm of the analytic trajectory."""
    tr, scans, imu = short_sequence
    od = oracle.Odometry(oracle.odom_default_config(map_voxel_size=0.5, preprocess_voxel_size=0.5))
    inserted = []
    poses = oracle.run_sequence(od, scans, imu, lambda i, o: inserted.append(o.info().last_inserted))
    assert np.array_equal(poses[0], np.eye(4)) and inserted[0] == 1       # Odometry.cpp:61
    G0 = tr.pose_world(scans[0][1][-1])
    for i in range(1, len(scans)):
        Gr = np.linalg.inv(G0) @ tr.pose_world(scans[i][1][-1])
        E = np.linalg.inv(Gr) @ poses[i]
        assert np.linalg.norm(E[:3, 3]) < 0.05
        assert np.arccos(np.clip(0.5 * (np.trace(E[:3, :3]) - 1), -1, 1)) < 0.01
    # the smooth start is slower than the 0.1 m / frame keyframe gate (LocalMap.cpp:132-147)
    assert inserted[1] == 0 and inserted[-1] == 1
    info = od.info()
    assert info.frames == len(scans) - 1 and info.n_states > 40 * (len(scans) - 2)
    # the states rolled back / replayed around every update end exactly on IMU stamps
    assert abs(od.last_state()["t"] * 400 - round(od.last_state()["t"] * 400)) < 1e-6
