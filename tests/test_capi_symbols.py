"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports
every symbol include/eskf_gpu.h declares, and fails loudly without a GPU (there
is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from eskf_lio_b200 import _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return _build.build()


def test_header_symbols_all_exported(built):
    hdr = open(os.path.join(ROOT, "include", "eskf_gpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|const char\*) (eskf_[a-z0-9_]+)\(", hdr, flags=re.M)))
    assert sorted(capi.SYMBOLS) == declared
    L = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(L, name), name
    assert L.eskf_abi_version() == 1


def test_no_torch_or_oracle_in_the_abi(built):
    hdr = open(os.path.join(ROOT, "include", "eskf_gpu.h")).read()
    assert "#include <torch" not in hdr and "at::" not in hdr and "<ATen" not in hdr
    for root, _, files in os.walk(os.path.join(ROOT, "eskf_lio_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "eskf_oracle.h" not in src and "liboracle" not in src, f


def test_fails_loudly_without_gpu(built):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.EskfError) as e:
        capi.Context(0)
    assert e.value.status == 3  # ESKF_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_host_driver_symbols_all_exported(built):
    """libeskf_host.so (include/eskf_host.h): Odometry + ErrorStateKF host classes."""
    from eskf_lio_b200 import odometry
    hdr = open(os.path.join(ROOT, "include", "eskf_host.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|const char\*) (eskf_[a-z0-9_]+)\(", hdr, flags=re.M)))
    assert sorted(odometry.SYMBOLS) == declared
    L = odometry.lib()
    for name in declared:
        assert hasattr(L, name), name
    cfg = odometry.default_config()
    assert cfg.map_voxel_size == 0.3 and cfg.max_iteration == 100 and cfg.imu_update_rate == 400.0
    if capi.device_count() == 0:
        with pytest.raises(RuntimeError) as e:
            odometry.Odometry(cfg)
        assert "no CPU fallback" in str(e.value)
