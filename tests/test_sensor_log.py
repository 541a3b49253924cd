"""the sensor wire format (include/ESKF_LIO/Subscriber.hpp:38-52, :80-103)
as a flat binary log: NumPy writer / reader, the C++ reader / writer behind include/eskf_host.h, and
(on the GPU box) a replay through the odometry that equals feeding the same records by hand."""
import ctypes as C
import os

import numpy as np
import pytest

from eskf_lio_b200 import sensor_log, synth as S


def small_log(n_frames=4, seed=5, decim=16):
    scans, imu = S.make_sequence(S.hall_scene(), S.hall_trajectory(), n_frames, seed=seed)
    scans = [(x[::decim].astype(np.float32), t[::decim].copy()) for x, t in scans]
    return scans, imu


def test_round_trip_numpy_and_cpp(tmp_path):
    from eskf_lio_b200 import odometry
    L = odometry.lib()      # loads without a GPU; nothing below launches a kernel
    scans, imu = small_log()
    path = str(tmp_path / "a.eskflog")
    sensor_log.write_sequence(path, scans, imu)
    recs = list(sensor_log.read(path))
    sweeps = [r for r in recs if r[0] == "lidar"]
    imus = [r for r in recs if r[0] == "imu"]
    assert len(sweeps) == len(scans) and len(imus) == len(imu)
    for (_, xyz, t), (x0, t0) in zip(sweeps, scans):
        np.testing.assert_array_equal(xyz, x0)
        np.testing.assert_array_equal(t, t0)
    np.testing.assert_array_equal(np.array([r[1] for r in imus]), imu[:, 0])
    np.testing.assert_array_equal(np.stack([r[2] for r in imus]), imu[:, 1:4])
    np.testing.assert_array_equal(np.stack([r[3] for r in imus]), imu[:, 4:7])
    # callback order: a sweep comes after every IMU sample up to its last point, before the next one
    order = [r[0] for r in recs]
    first_sweep = order.index("lidar")
    assert all(r[1] <= scans[0][1][-1] for r in recs[:first_sweep]) and recs[first_sweep + 1][1] > scans[0][1][-1]
    # the C++ reader sees the same file
    counts = (C.c_uint64 * 3)()
    stamps = (C.c_double * 2)()
    assert L.eskf_log_summary(path.encode(), counts, stamps) == 0
    assert list(counts) == [len(imu), len(scans), sum(len(t) for _, t in scans)]
    assert stamps[0] == min(imu[0, 0], scans[0][1][0]) and stamps[1] == max(imu[-1, 0], scans[-1][1][-1])
    # the C++ writer produces the identical bytes
    path2 = str(tmp_path / "b.eskflog")
    w = C.c_void_p()
    assert L.eskf_log_writer_open(path2.encode(), C.byref(w)) == 0
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    for r in recs:
        if r[0] == "imu":
            g, a = np.ascontiguousarray(r[2]), np.ascontiguousarray(r[3])
            assert L.eskf_log_writer_imu(w, C.c_double(r[1]), g.ctypes.data_as(dp), a.ctypes.data_as(dp)) == 0
        else:
            x, t = np.ascontiguousarray(r[1], dtype=np.float32), np.ascontiguousarray(r[2])
            assert L.eskf_log_writer_lidar(w, x.ctypes.data_as(fp), t.ctypes.data_as(dp), C.c_size_t(len(t))) == 0
    assert L.eskf_log_writer_close(w) == 0
    assert open(path, "rb").read() == open(path2, "rb").read()


def test_malformed_logs_are_refused(tmp_path):
    from eskf_lio_b200 import odometry
    L = odometry.lib()
    scans, imu = small_log(n_frames=2)
    path = str(tmp_path / "a.eskflog")
    sensor_log.write_sequence(path, scans, imu)
    blob = open(path, "rb").read()
    counts, stamps = (C.c_uint64 * 3)(), (C.c_double * 2)()
    bad = str(tmp_path / "bad.eskflog")
    for data in (b"NOTALOG!" + blob[8:], blob[:len(blob) - 7], blob[:16] + b"\x09\x00\x00\x00\x01\x00\x00\x00"):
        open(bad, "wb").write(data)
        assert L.eskf_log_summary(bad.encode(), counts, stamps) != 0
        assert L.eskf_host_last_error()
        with pytest.raises(ValueError):
            list(sensor_log.read(bad))
    assert L.eskf_log_summary(str(tmp_path / "missing").encode(), counts, stamps) != 0
    # an empty log (header only) is fine
    open(bad, "wb").write(blob[:16])
    assert L.eskf_log_summary(bad.encode(), counts, stamps) == 0 and list(counts) == [0, 0, 0]
    assert list(sensor_log.read(bad)) == []


@pytest.mark.gpu
@pytest.mark.parametrize("device_resident", [1, 0])
def test_replay_equals_feeding_by_hand(tmp_path, device_resident):
    """eskf_odom_replay_log over the file == odometry.run_sequence over the arrays: bit-identical poses."""
    from eskf_lio_b200 import odometry
    scans, imu = S.make_sequence(S.hall_scene(), S.hall_trajectory(), 12, seed=17)
    scans = [(x.astype(np.float32), t) for x, t in scans]
    path = str(tmp_path / "seq.eskflog")
    sensor_log.write_sequence(path, scans, imu)
    cfg = dict(map_voxel_size=0.5, preprocess_voxel_size=0.5, device_resident=device_resident)
    od = odometry.Odometry(odometry.default_config(**cfg))
    want = odometry.run_sequence(od, scans, imu)
    od.close()
    od = odometry.Odometry(odometry.default_config(**cfg))
    got = od.replay_log(path)
    assert len(got) == len(want) == 12 and od.info().frames == 11
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    od.close()
