"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libeskf_gpu.so")
HOST_DIR = os.path.join(HERE, "host")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libeskf_host.so")
INCLUDE_DIR = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    root = os.path.dirname(HERE)
    return sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(root, "include", "eskf_gpu.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def host_is_stale() -> bool:
    if not os.path.exists(HOST_LIB_PATH):
        return True
    t = os.path.getmtime(HOST_LIB_PATH)
    deps = glob.glob(os.path.join(HOST_DIR, "*.cpp")) + glob.glob(os.path.join(HOST_DIR, "ESKF_LIO", "*.hpp")) + \
        glob.glob(os.path.join(INCLUDE_DIR, "*.h")) + [LIB_PATH]
    return any(os.path.getmtime(p) > t for p in deps)


def build_host(force: bool = False) -> str:
    """Compile eskf_lio_b200/lib/libeskf_host.so: the ROS-free Odometry / ErrorStateKF host
    classes behind include/eskf_host.h (plain g++; calls the GPU only through libeskf_gpu.so)."""
    if not force and not host_is_stale():
        return HOST_LIB_PATH
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-fPIC", "-shared",
           f"-I{INCLUDE_DIR}", f"-I{HOST_DIR}", "-o", HOST_LIB_PATH,
           os.path.join(HOST_DIR, "odometry_capi.cpp"), f"-L{LIB_DIR}", "-leskf_gpu", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return HOST_LIB_PATH


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile eskf_lio_b200/lib/libeskf_gpu.so with nvcc for sm_100a (+ the host driver library)."""
    path = _build_gpu(force, verbose)
    build_host(force)
    return path


def _build_gpu(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:" + env.get("PATH", "")  # host compiler: the distro g++ (has libgomp etc.)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH
