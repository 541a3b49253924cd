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


def test_ctypes_structs_match_the_c_headers(tmp_path):
    """The ctypes mirrors of the structs that cross the C ABI have the size and field offsets the
    C compiler gives the headers' definitions."""
    import subprocess
    from eskf_lio_b200 import odometry
    pairs = [("eskf_odom_config", odometry.OdomConfig, "eskf_host.h"),
             ("eskf_odom_info", odometry.OdomInfo, "eskf_host.h"),
             ("eskf_state", capi.State, "eskf_gpu.h"),
             ("eskf_icp_params", capi.IcpParams, "eskf_gpu.h"),
             ("eskf_align_info", capi.AlignInfo, "eskf_gpu.h")]
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "eskf_gpu.h"', '#include "eskf_host.h"',
             'int main(void) {']
    for cname, cls, _ in pairs:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, *_ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    for (cname, cls, _), line in zip(pairs, out):
        nums = [int(x) for x in line.split()[1:]]
        assert nums[0] == ctypes.sizeof(cls), cname
        assert nums[1:] == [getattr(cls, f[0]).offset for f in cls._fields_], cname


def test_align_batch_argument_checks_without_gpu(built):
    """eskf_align_batch validates its context list and returns for an empty batch before it touches
    any context (the array marshalling of the ctypes binding is exercised here, the compute in -m gpu)."""
    class Fake:
        def __init__(self, v):
            self._h = ctypes.c_void_p(v)
    assert capi.align_batch([Fake(0x1000), Fake(0x2000)], [], [], []) == []
    with pytest.raises(capi.EskfError) as e:
        capi.align_batch([Fake(0x1000), Fake(0x1000)], [], [], [])
    assert "distinct" in str(e.value)
    with pytest.raises(ValueError):
        capi.align_batch([Fake(0x1000)], [Fake(1)], [], [])


def test_stamps_sorted_host_helper(built):
    """eskf_stamps_sorted (host-only): the precondition of the reference's deskew
    (src/CloudPreprocessor.cpp:33) as a question; equal stamps count as sorted, block boundaries
    of the vectorised loop (4096) and tiny inputs are covered."""
    import numpy as np
    t = np.arange(20000) * 1.5e-6
    assert capi.stamps_sorted(t) and capi.stamps_sorted(t[:1]) and capi.stamps_sorted(t[:0])
    assert capi.stamps_sorted(np.repeat(t[:100], 3))
    for k in (1, 2, 4095, 4096, 4097, 8192, 19999):
        u = t.copy()
        u[k] = u[k - 1] - 1e-9
        assert not capi.stamps_sorted(u), k
