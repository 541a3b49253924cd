"""The C++ drop-in classes (eskf_lio_b200/host/ESKF_LIO/*.hpp: ICP, LocalMap,
CloudPreprocessor with the reference's names and call order) driven like
Odometry::run (src/Odometry.cpp:55-87) and checked against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from eskf_lio_b200 import _build, synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_classes.cpp")


def _compile(out, syntax_only=False, src=SRC):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra",
           f"-I{ROOT}/include", f"-I{ROOT}/eskf_lio_b200/host"]
    if syntax_only:
        cmd += ["-fsyntax-only", src]
    else:
        cmd += ["-o", out, src, f"-L{_build.LIB_DIR}", "-leskf_gpu", f"-Wl,-rpath,{_build.LIB_DIR}"]
    subprocess.check_call(cmd)


def test_host_classes_compile():
    """CPU: the headers are self-contained C++17 over include/eskf_gpu.h only."""
    _compile(None, syntax_only=True)
    _compile(None, syntax_only=True, src=os.path.join(ROOT, "eskf_lio_b200", "host", "odometry_capi.cpp"))
    for h in ("Types.hpp", "LocalMap.hpp", "Registration.hpp", "CloudPreprocessor.hpp",
              "ErrorStateKF.hpp", "Odometry.hpp", "SynchronizedQueue.hpp", "GpuContext.hpp"):
        src = open(os.path.join(ROOT, "eskf_lio_b200", "host", "ESKF_LIO", h)).read()
        assert "Eigen/" not in src and "open3d" not in src.replace("open3d::", "") and "oracle" not in src


def test_host_eskf_process_matches_oracle(tmp_path, oracle):
    """CPU: the product's host ErrorStateKF::process (src/ErrorStateKF.cpp:76-113) against the
    oracle's restatement on the same IMU stream (no GPU call is involved in propagation)."""
    _build.build()
    exe = str(tmp_path / "test_eskf_host")
    _compile(exe, src=os.path.join(ROOT, "tests", "cpp", "test_eskf_host.cpp"))
    tr = S.hall_trajectory()
    rng = np.random.default_rng(9)
    rows = []
    for k in range(1, 301):
        g, a = tr.imu(k / 400.0, rng, 0.005, 0.02)
        rows.append([k / 400.0, *g, *a])
    rows.insert(100, [0.1, 0, 0, 0, 0, 0, 0])   # a stale sample: dt < 0 -> ignored (:80-82)
    rows = np.array(rows)
    path = str(tmp_path / "imu.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(rows)))
        f.write(rows.tobytes())
    out = subprocess.check_output([exe, path], text=True).strip().splitlines()
    od = oracle.Odometry()
    for r in rows:
        od.kf_process(r[0], r[1:4], r[4:7])
    st = od.last_state(with_P=True)
    assert int(out[0].split()[1]) == od.info().n_states == 301
    got = np.array([float(v) for v in out[1].split()[1:]])
    want = np.concatenate([[st["t"]], st["p"], st["v"], st["q"]])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    P = np.array([float(v) for v in out[2].split()[1:]]).reshape(18, 18)
    assert np.linalg.norm(P - st["P"]) / np.linalg.norm(st["P"]) < 1e-12
    from scipy.spatial.transform import Rotation
    ax = np.array([0.2, -0.5, 0.84])
    ax /= np.linalg.norm(ax)
    for line, ang in zip(out[3:5], (0.3, 2.9)):
        rv = np.array([float(v) for v in line.split()[1:]])
        np.testing.assert_allclose(rv, Rotation.from_rotvec(ang * ax).as_rotvec(), atol=1e-12)


@pytest.mark.gpu
def test_host_classes_match_oracle(tmp_path, oracle):
    _build.build()
    exe = str(tmp_path / "test_host")
    _compile(exe)
    rng = np.random.default_rng(21)
    scene = S.hall_scene()
    poses = S.arc_trajectory(4)
    T_il = S.default_T_il()
    scans = [S.make_scan(scene, T, rng) for T in poses]
    guess = poses[3] @ S.perturbation()
    path = str(tmp_path / "in.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(scans)))
        for k, (xyz, t) in enumerate(scans):
            xyz, t = xyz[::2], t[::2]
            f.write(struct.pack("<Q", len(xyz)))
            f.write(np.ascontiguousarray(xyz).tobytes())
            f.write(np.ascontiguousarray(t).tobytes())
            f.write(np.ascontiguousarray(guess if k == 3 else poses[k]).tobytes())
    pcd_path, traj_path = str(tmp_path / "map_cloud.pcd"), str(tmp_path / "trajectory.json")
    out = subprocess.check_output([exe, path, pcd_path, traj_path], text=True)
    lines = out.strip().splitlines()
    # oracle replay of the same call sequence
    om = oracle.Map(0.5, 1000)
    kept, maps, first = [], [], []
    for k, (xyz, t) in enumerate(scans):
        p, c, _ = oracle.preprocess(xyz[::2], t[::2], T_il, None, 0.5)
        kept.append(len(p))
        T = poses[k]
        if k == 3:
            r = om.align(p, c, guess)
            T = r["T"]
        _, _, pw, _ = om.update(p, c, T, initialize=True)
        maps.append(om.size())
        first.append(pw[0])
    got_kept = [int(l.split()[3]) for l in lines if l.startswith("scan")]
    assert got_kept == kept
    al = [l for l in lines if l.startswith("align")][0].split()
    assert int(al[2]) == r["iterations"]
    Tg = np.array([float(v) for v in al[4:20]]).reshape(4, 4)
    E = np.linalg.inv(r["T"]) @ Tg
    assert np.linalg.norm(E[:3, 3]) < 1e-5
    assert np.arccos(np.clip(0.5 * (np.trace(E[:3, :3]) - 1), -1, 1)) < 1e-5
    got_maps = [int(l.split()[1]) for l in lines if l.startswith("map")]
    assert got_maps[:3] == maps[:3]          # identical poses -> identical occupancy
    assert abs(got_maps[3] - maps[3]) <= 2   # pose differs by ~1e-9 m: a boundary voxel may flip
    for k in range(3):
        w = np.array([float(v) for v in [l for l in lines if l.startswith("map")][k].split()[3:6]])
        np.testing.assert_array_equal(w, first[k])  # caller's cloud left in the world frame, bit-exact
    assert "ICP not converged" not in out
    # LocalMap::save (src/LocalMap.cpp:156-167): one point per voxel + the pose of every frame
    import json
    traj = json.load(open(traj_path))
    assert traj["class_name"] == "PinholeCameraTrajectory" and len(traj["parameters"]) == len(scans)
    E = np.array(traj["parameters"][0]["extrinsic"]).reshape(4, 4).T   # column-major like Open3D
    np.testing.assert_allclose(E, poses[0], atol=1e-15)
    body = open(pcd_path).read().split("DATA ascii\n")[1].strip().splitlines()
    assert len(body) == got_maps[-1]
