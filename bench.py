#!/usr/bin/env python
"""bench.py — headline benchmark of the ESKF_LIO hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline (every N, so that the driver's 1 -> 8 scaling column is about the
path that shards): BASELINE.json configs[2], the dense registration — a
2 M-point source against the 10 M-point voxel map at 0.1 m, the source split
by point range over the N GPUs, the map replicated, the 28 H/b sums of every
Gauss-Newton iteration exchanged over NVLink inside the persistent kernel
(eskf_align_cloud_p2p; reference: src/Registration.cpp:60-76, the reduction
under `omp critical`).  A step = one registration of fixed 10 GN iterations;
metric = Mpts/s per GN iteration (BASELINE.json `metric`), strong scaling.

  value     shards resident in HBM; CUDA events on the context's stream around
            the K registrations, max over ranks.
  e2e       the same registrations through the host-buffer C ABI: every step
            uploads its shard from pinned host memory (eskf_cloud_upload),
            registers it and reads the pose back; wall clock, max over ranks.
  roofline  align_kernel, timed alone (CUDA events, median of 5 launches).
  parity    in the same run: every rank holds the bit-identical pose; N > 1:
            rank 0 also registers the unsharded source — iteration count and
            per-iteration correspondence counts equal, pose within 1e-5;
            N = 1: the CPU oracle on a bounded sample of the source.
  cpu_baseline / --impl reference
            the CPU oracle (dependency-free restatement of the reference's
            OpenMP path; the reference itself cannot be built here) on the host
            cores: the same map, a bounded sample of the same source.

Sub-records: `frame` = BASELINE.json configs[1], the 64k-point odometry
sequence (ms/frame, device-timed and end to end, p99 / max, against the CPU
oracle; one replica per GPU at N > 1: a frame does not shard); `single_scan` =
configs[0]; `weak` = 2 M points per GPU; `nccl_baseline` = the same sharded
registration with an NCCL all-reduce between two launches per iteration;
`batch` = configs[4], independent scan / map pairs, one shard per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from eskf_lio_b200 import synth as S  # noqa: E402

ALG_BYTES_PER_POINT_ITER = 136  # 24 pos rd + 24 pos wr + 24 src cov + 64 voxel slot
DENSE_SRC = int(os.environ.get("ESKF_BENCH_DENSE_SRC", "2000000"))  # (the contract test shrinks these)
DENSE_MAP = int(os.environ.get("ESKF_BENCH_DENSE_MAP", "10000000"))
DENSE_VOXEL = 0.1
DENSE_ITERS = 10
DENSE_SOURCES = 3            # distinct source clouds the steps rotate through
CPU_SAMPLE = int(os.environ.get("ESKF_BENCH_CPU_SAMPLE", "200000"))  # source points per CPU-oracle step
GUESS = dict(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
# configs[1]: reference defaults (config/hilti_config.yaml): 0.3 m voxels, cap 1000, ICP 100 / 1e-6 / 0.9999
VOXEL = 0.3
# untimed lead-in: frame 0 initialises the map, the trajectory starts from rest (the keyframe gate of
# LocalMap.cpp:132-147 only opens at ~1 m/s), so ~35 frames put >= 20 scans into the local map
LEAD_IN = int(os.environ.get("ESKF_BENCH_LEAD_IN", "35"))
FRAME_WARMUP = 5
FRAME_STEPS = int(os.environ.get("ESKF_BENCH_FRAMES", "70"))  # timed frames: past frame 100, where the
#                                                               10 s eviction sweep of LocalMap.cpp:60-72 falls
CACHE_DIR = os.environ.get("ESKF_BENCH_CACHE", "/tmp/eskf_lio_b200_cache")
MAP_HINT = int(os.environ.get("ESKF_BENCH_MAP_HINT", "0"))  # 0: Config::local_map.capacity_hint's default


def make_log(n_frames, seed):
    """The synthetic sensor log of configs[1]: 400x30x10 m corridor,
    10 Hz motion-distorted 32x2000 sweeps + 400 Hz IMU from an analytic trajectory (~1.2 m/s).
    Cached on disk so the reference arm (a separate process) replays the identical log."""
    path = os.path.join(CACHE_DIR, f"corridor_v1_{n_frames}_{seed}.npz")
    if os.path.exists(path):
        try:
            z = np.load(path)
            n = z["n"]
            xyz, t = z["xyz"], z["t"]
            off = np.concatenate([[0], np.cumsum(n)])
            scans = [(xyz[off[i]:off[i + 1]].astype(np.float64), t[off[i]:off[i + 1]].copy())
                     for i in range(len(n))]
            return scans, z["imu"]
        except Exception:
            pass
    scans, imu = S.make_sequence(S.corridor_scene(), S.corridor_trajectory(), n_frames, seed, chunk=100)
    try:
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, n=np.array([len(x) for x, _ in scans]),
                 xyz=np.concatenate([x for x, _ in scans]).astype(np.float32),
                 t=np.concatenate([t for _, t in scans]), imu=imu)
        os.replace(tmp, path)
    except Exception:
        pass
    return scans, imu


def odom_overrides():
    return dict(map_voxel_size=VOXEL, preprocess_voxel_size=VOXEL)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", p
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pose_delta(A, B):
    E = np.linalg.inv(A) @ B
    R = E[:3, :3]
    sin = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.linalg.norm(E[:3, 3])), float(np.arctan2(sin, 0.5 * (np.trace(R) - 1.0)))


# ------------------------------------------------------------ dense workload
def dense_map_chunks():
    """The 10 M map points of configs[2] (seed 44), in insert-sized chunks; the generator's state
    then yields the source clouds, so every process draws the identical data."""
    rng = np.random.default_rng(44)
    scene = S.block_scene()
    left = DENSE_MAP
    chunks = []
    while left > 0:
        n = min(2_500_000, left)
        chunks.append(n)
        left -= n
    return rng, scene, chunks


def dense_workload_config(world):
    return {"workload": f"BASELINE.json configs[2]: dense undownsampled registration, {DENSE_SRC}-point source vs "
                        f"{DENSE_MAP}-point voxel map at {DENSE_VOXEL} m voxels (200x200x20 m block scene, seed 44), "
                        f"source sharded by point range over {world} GPU(s), map replicated, per-iteration H/b "
                        "exchange over NVLink inside the persistent kernel; a step = one registration of "
                        f"{DENSE_ITERS} fixed Gauss-Newton iterations; guess = 0.03 m / 0.3 deg perturbation",
            "n_source_points": DENSE_SRC, "n_map_points": DENSE_MAP, "voxel_size": DENSE_VOXEL,
            "gn_iterations_per_step": DENSE_ITERS, "neighbour_mode": 1, "parallelism": f"point-range x{world}",
            "l2": f"steps rotate through {DENSE_SOURCES} distinct source clouds; per-GPU inputs "
                  "(shards 192 MB / N each + 5.9 GB map) exceed the 126 MB L2 up to N = 4; at N = 8 the three 24 MB "
                  "shards + the touched records fit, as they would in production (GN iterations re-read them by design)"}


class DenseOracle:
    """CPU oracle leg of the dense workload: the same map, a bounded sample of the source."""

    def __init__(self, threads=None):
        import oracle as O
        O.build()
        O.set_num_threads(threads or len(os.sched_getaffinity(0)))
        self.O = O
        rng, scene, chunks = dense_map_chunks()
        self.map = O.Map(DENSE_VOXEL, 1000)
        t0 = time.perf_counter()
        for n in chunks:
            p, c = S.dense_cloud(scene, n, rng)
            self.map.update(p, c, np.eye(4), initialize=True)
        self.build_s = time.perf_counter() - t0
        self.sources = [S.dense_cloud(scene, DENSE_SRC, rng) for _ in range(DENSE_SOURCES)]
        self.guess = S.perturbation(**GUESS)

    def step(self, k):
        p, c = self.sources[k % DENSE_SOURCES]
        m = min(CPU_SAMPLE, len(p))
        t0 = time.perf_counter()
        r = self.map.align(p[:m], c[:m], self.guess, max_iteration=DENSE_ITERS,
                           translation_sq_threshold=0.0, cosine_threshold=2.0)
        return time.perf_counter() - t0, r, m

    def run(self, steps, warmup):
        for k in range(warmup):
            self.step(k)
        total, pts = 0.0, 0
        for k in range(steps):
            dt, r, m = self.step(warmup + k)
            total += dt
            pts += m * r["iterations"]
        return {"mpts_per_s": pts / total / 1e6, "seconds": total, "steps": steps, "cores": self.O.num_threads(),
                "ms_per_step": 1e3 * total / max(steps, 1), "sample_points": min(CPU_SAMPLE, DENSE_SRC),
                "map_build_s": self.build_s, "map_voxels": self.map.size()}


def cpu_sample_text(r):
    return (f"{r['steps']} registrations of the first {r['sample_points']} points of the source against the same "
            f"{DENSE_MAP}-point map, {DENSE_ITERS} fixed GN iterations each, all host threads (oracle = "
            "dependency-free restatement of the reference's OpenMP path, which fuses the reference's two passes "
            "— correspondenceMatching then computeTransform — into one; the reference itself needs "
            "Eigen/Open3D/yaml-cpp/rclcpp and cannot be built here)")


def impl_reference(args, rank, world):
    if rank != 0:
        return
    orc = DenseOracle()
    r = orc.run(args.steps, min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": "mpts_per_s_per_gn_iteration", "value": r["mpts_per_s"], "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dense_workload_config(args.gpus),
        "cpu_baseline": {"value": r["mpts_per_s"], "unit": "Mpts/s", "cores": r["cores"], "kind": "port",
                         "sample": cpu_sample_text(r), "map_build_s": r["map_build_s"]},
        "e2e": {"value": r["mpts_per_s"], "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


SAME_DEVICE = os.environ.get("ESKF_BENCH_SAME_DEVICE", "") not in ("", "0")  # tests: every rank on cuda:0, gloo


class Dist:
    """torch.distributed plumbing (barrier, max / min over ranks); a no-op at world 1."""

    def __init__(self, world, local_rank):
        import torch
        self.torch = torch
        self.world = world
        self.local = 0 if SAME_DEVICE else local_rank
        self.dist = None
        self.nccl = False
        torch.cuda.set_device(self.local)
        if world > 1:
            import torch.distributed as dist
            if SAME_DEVICE:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                self.nccl = True
            self.dist = dist
        self.dev = f"cuda:{self.local}" if (self.nccl or world == 1) else "cpu"

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, vals, op):
        if self.dist is None:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return [float(x) for x in t]

    def max(self, *vals):
        return self._reduce(vals, self.dist.ReduceOp.MAX if self.dist else None)

    def min(self, *vals):
        return self._reduce(vals, self.dist.ReduceOp.MIN if self.dist else None)

    def gather(self, val):
        if self.dist is None:
            return [float(val)]
        t = self.torch.tensor([float(val)], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o[0]) for o in out]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def pinned_copy(capi, arr):
    """A page-locked host copy of `arr` (what a caller that cares about PCIe hands the C ABI)."""
    import ctypes as C
    a = np.ascontiguousarray(arr, dtype=np.float64)
    p = C.c_void_p()
    capi.check(capi.lib().eskf_host_alloc(C.c_size_t(max(a.nbytes, 8)), C.byref(p)))
    C.memmove(p, a.ctypes.data, a.nbytes)
    return p


def dense_leg(args, capi, sharded, D, rank, world, local_rank):
    """configs[2] on `world` GPUs.  Returns the headline numbers (all ranks) and rank 0's sub-records."""
    ctx = capi.Context(local_rank)
    rng, scene, chunks = dense_map_chunks()
    gmap = capi.Map(ctx, DENSE_VOXEL, 1000, max(1 << 16, int(0.9 * DENSE_MAP)))
    for n in chunks:
        p, c = S.dense_cloud(scene, n, rng)
        gmap.insert(p, c, np.eye(4))
    n_vox = gmap.size()
    b, e = sharded.shard_range(DENSE_SRC, rank, world)
    clouds, pinned, full0 = [], [], None
    for k in range(DENSE_SOURCES):
        p, c = S.dense_cloud(scene, DENSE_SRC, rng)
        if k == 0 and rank == 0:
            full0 = (p, c)
        clouds.append(capi.Cloud(ctx, max(e - b, 64)).upload(p[b:e], c[b:e]))
        pinned.append((pinned_copy(capi, p[b:e]), pinned_copy(capi, c[b:e].reshape(-1, 9))))
    del p, c
    guess = S.perturbation(**GUESS)
    comm = sharded.make_comm(ctx) if world > 1 else capi.Comm(ctx, 0, 1)
    e2e_cloud = capi.Cloud(ctx, max(e - b, 64))
    ctx.sync()

    step, _ = gmap.p2p_stepper(clouds, guess, comm, DENSE_ITERS)  # (arguments marshalled once)

    def step_e2e(k):
        px, pc = pinned[k % DENSE_SOURCES]
        e2e_cloud.upload_ptr(px.value, pc.value, e - b)
        return gmap.align_cloud_p2p(e2e_cloud, guess, comm, fixed_iterations=DENSE_ITERS)

    # ---- value: shards resident in HBM, CUDA events around the K registrations
    for k in range(args.warmup):
        step(k)
    D.barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for k in range(args.steps):
        step(args.warmup + k)
    ms_dev = ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    D.barrier()
    # ---- e2e: every step uploads its shard from pinned host memory and reads the pose back
    for k in range(args.warmup):
        step_e2e(k)
    D.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(args.warmup + k)
    ctx.sync()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    D.barrier()
    clocks = sampler.stop() if sampler else None
    ms_dev, ms_e2e = D.max(ms_dev, ms_e2e)
    out = {"ms_dev": ms_dev, "ms_e2e": ms_e2e, "launches": launches, "clocks": clocks, "n_vox": n_vox,
           "shard_points": e - b, "h2d": (e - b) * 96, "d2h": 16 * 8 + 64}

    # ---- the kernel alone (roofline): single launches, CUDA events, median of 5
    times, ncorr = [], []
    for k in range(5):
        D.barrier()
        ctx.timer_start()
        r = gmap.align_cloud_p2p(clouds[0], guess, comm, fixed_iterations=DENSE_ITERS, trace=True)
        times.append(ctx.timer_stop())
        ncorr = [int(v) for v in r["ncorr"]]
    out["ms_launch"] = D.max(float(np.median(times)))[0]
    out["ncorr_fixed"] = ncorr
    # which of the two large-cloud loop shapes this context's one-off timing kept (eskf_gpu.h "align_autotune")
    tb = ctx.get_option("align_tuned_block")
    out["kernel_shape"] = {512: "4-deep load rotation, 512 threads x 1 CTA/SM, 8-bit probe filter",
                           256: "3-stage pipeline, 256 threads x 3 CTAs/SM, 16-bit tags",
                           257: "3-stage pipeline, 256 threads x 3 CTAs/SM, 8-bit probe filter",
                           769: "3-stage pipeline, 768 threads x 1 CTA/SM, 16-bit tags"}.get(tb, f"{tb} threads")

    # ---- parity, in the same run
    D.barrier()
    r_conv = gmap.align_cloud_p2p(clouds[0], guess, comm, trace=True)
    lo = D.min(*r_conv["T"].ravel())
    hi = D.max(*r_conv["T"].ravel())
    parity = {"identical_pose_on_all_ranks": bool(lo == hi), "converged_iterations": r_conv["iterations"],
              "converged": bool(r_conv["converged"])}
    if world > 1 and D.nccl:
        # the same registration with an NCCL all-reduce between two launches per iteration (baseline)
        cb = sharded.TorchAllReduce()
        ts = []
        for k in range(2 + 3):
            D.barrier()
            ctx.timer_start()
            r_nccl = gmap.align_cloud_sharded(clouds[0], guess, cb, fixed_iterations=DENSE_ITERS)
            ts.append(ctx.timer_stop())
        out["nccl_ms_launch"] = D.max(float(np.median(ts[2:])))[0]
        r_p2p = gmap.align_cloud_p2p(clouds[0], guess, comm, fixed_iterations=DENSE_ITERS)
        parity["nccl_vs_fused_pose_delta"] = pose_delta(r_nccl["T"], r_p2p["T"])
    if rank == 0 and world > 1:
        full = capi.Cloud(ctx, DENSE_SRC).upload(*full0)
        ref = gmap.align_cloud(full, guess, trace=True)
        nit = min(ref["iterations"], r_conv["iterations"])
        dt, dr = pose_delta(ref["T"], r_conv["T"])
        parity.update({
            "unsharded_iterations": ref["iterations"],
            "ncorr_equal_to_unsharded": bool(np.array_equal(ref["ncorr"][:nit], r_conv["ncorr"][:nit])),
            "pose_vs_unsharded": {"m": dt, "rad": dr},
            "H_rel_vs_unsharded_max": float(max(np.linalg.norm(r_conv["H"][k] - ref["H"][k]) /
                                                np.linalg.norm(ref["H"][k]) for k in range(nit)))})
        assert ref["iterations"] == r_conv["iterations"], "sharded / unsharded iteration counts differ"
        assert parity["ncorr_equal_to_unsharded"], "sharded / unsharded correspondence counts differ"
        assert dt < 1e-5 and dr < 1e-5, ("sharded pose off the unsharded one", dt, dr)
        # strong-scaling denominator measured in this very run: the unsharded registration on rank 0's GPU
        ts = []
        for k in range(2 + 3):
            ctx.timer_start()
            gmap.align_cloud_fixed(full, guess, DENSE_ITERS)
            ts.append(ctx.timer_stop())
        out["unsharded_ms_launch_rank0"] = float(np.median(ts[2:]))
        del full
    assert parity["identical_pose_on_all_ranks"], "ranks ended on different poses"
    D.barrier()
    out["parity"] = parity

    # ---- weak scaling: DENSE_SRC points on every GPU
    if world > 1 and not args.no_weak:
        p, c = S.dense_cloud(scene, DENSE_SRC, np.random.default_rng(4400 + rank))
        wc = capi.Cloud(ctx, DENSE_SRC).upload(p, c)
        ts = []
        for k in range(2 + 5):
            D.barrier()
            ctx.timer_start()
            gmap.align_cloud_p2p(wc, guess, comm, fixed_iterations=DENSE_ITERS)
            ts.append(ctx.timer_stop())
        out["weak_ms_launch"] = D.max(float(np.median(ts[2:])))[0]
        del wc, p, c
    D.barrier()
    out["sample_gpu"] = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the registration the CPU oracle is held against below: the bounded sample, run to convergence
        m = min(CPU_SAMPLE, DENSE_SRC)
        sc = capi.Cloud(ctx, m).upload(full0[0][:m], full0[1][:m])
        out["sample_gpu"] = gmap.align_cloud(sc, guess, trace=True)
        del sc
    comm.close()
    for px, pc in pinned:
        capi.lib().eskf_host_free(px)
        capi.lib().eskf_host_free(pc)
    del clouds, e2e_cloud, gmap
    ctx.close()
    out["full0"] = full0
    return out


# ------------------------------------------------------------ frame workload
def replay(odom, scans, imu, feed, first_timed, on_timed_start=None):
    """Deliver the log the way the two sensor callbacks would (IMU samples in time order, a sweep
    once its last point is measured, one spin per delivery) and time, per frame from `first_timed`
    on, the delivery of the sweep + the spin that consumes it (wall clock)."""
    import gc
    k = 0
    n_imu = imu.shape[0]
    poses, iters, walls = [], [], []
    gc.collect()
    gc.disable()  # no collector pauses inside the wall-clock legs (re-enabled below)
    for i, (xyz, t) in enumerate(scans):
        if i == first_timed and on_timed_start:
            on_timed_start()
        end = t[-1]
        while k < n_imu and imu[k, 0] <= end:
            odom.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
            odom.spin_once()
        # the first IMU sample past the sweep end makes the frame eligible (Odometry.cpp:65-69)
        if k < n_imu:
            odom.feed_imu(imu[k, 0], imu[k, 1:4], imu[k, 4:7])
            k += 1
        t0 = time.perf_counter()
        feed(i)
        done = odom.spin_once()
        pose = odom.pose()
        t1 = time.perf_counter()
        if not done:
            raise RuntimeError("frame not consumed: IMU stream too short")
        if i >= first_timed:
            walls.append(1e3 * (t1 - t0))
            iters.append(odom.info().last_iterations)
        poses.append(pose)
    gc.enable()
    return walls, poses, iters


def run_frame_oracle(scans, imu, first_timed, threads=None):
    """The CPU oracle (restatement of the reference's OpenMP path incl. its ErrorStateKF and the call
    order of Odometry::run) on the same log; ms/frame = the reference's own three stage timers
    (src/Odometry.cpp:73-87), summed over the frames from `first_timed` on."""
    import oracle as O
    O.build()
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))  # (torchrun exports OMP_NUM_THREADS=1)
    od = O.Odometry(O.odom_default_config(**odom_overrides()))
    base = [None]

    def snap():
        inf = od.info()
        return np.array(inf.stage_avg_ms) * inf.frames, inf.frames

    _, poses, iters = replay(od, scans, imu, lambda i: od.feed_lidar(scans[i][0], scans[i][1]),
                             first_timed, lambda: base.__setitem__(0, snap()))
    s1, f1 = snap()
    s0, f0 = base[0]
    n = f1 - f0
    stage = (s1 - s0) / max(n, 1)
    return {"ms_per_frame": float(stage.sum()), "frames": int(n), "cores": O.num_threads(),
            "stage_ms": {"preprocess": float(stage[0]), "filter_update": float(stage[1]),
                         "map_update": float(stage[2])},
            "gn_iterations_mean": float(np.mean(iters)) if iters else 0.0, "poses": poses}


def frame_workload_config():
    return {"workload": "BASELINE.json configs[1]: synthetic odometry sequence, 10 Hz 32-beam x 2000-col "
                        "(64k-pt) motion-distorted sweeps + 400 Hz IMU in a 400x30x10 m corridor (~1.2 m/s); "
                        "per frame: preprocess (T_il, deskew against the filter states, 0.3 m downsample, "
                        "30-NN covariances) -> host ErrorStateKF::update {VGICP align from the IMU-predicted "
                        "pose} -> LocalMap insert/evict; ms/frame = the three stages the reference times "
                        "(src/Odometry.cpp:73-87); the 400 Hz IMU propagation between frames is host work in "
                        "both arms and reported separately",
            "voxel_size": VOXEL, "lead_in_frames": LEAD_IN, "warmup_frames": FRAME_WARMUP,
            "timed_frames": FRAME_STEPS,
            "registration": "max_iteration=100, translation_sq_threshold=1e-6, cosine_threshold=0.9999, "
                            "1-neighbour (config/hilti_config.yaml)",
            "l2": "frame working set (~4 MB) is L2-resident by nature and every frame is a different sweep"}


def gpu_sequence(capi, odometry, device, scans, imu, first_timed, mode):
    """mode 'resident': every raw sweep is uploaded to HBM before the timed region (the frames are
    fed as device clouds); 'e2e': sweeps sit in pinned host memory in the float32 wire format and
    are copied to the device when they are delivered, the pose is read back every frame; 'host':
    like 'e2e' but through the plain drop-in classes (Config::device_resident = false): every one
    of the three calls takes and returns host vectors like the reference's, so the downsampled
    cloud crosses PCIe three times per frame."""
    import ctypes as C
    extra = {"map_capacity_hint": MAP_HINT} if MAP_HINT else {}
    od = odometry.Odometry(odometry.default_config(device_resident=0 if mode == "host" else 1,
                                                   **odom_overrides(), **extra), device)
    ctx = od.context()
    keep = []
    if mode == "resident":
        clouds = [capi.Cloud(ctx, len(x)).upload_f32(x) for x, _ in scans]
        ctx.sync()

        def feed(i):
            od.feed_lidar_cloud(clouds[i], scans[i][1])
    else:
        pinned = []
        for x, t in scans:
            x32 = np.ascontiguousarray(x, dtype=np.float32)
            px, pt = C.c_void_p(), C.c_void_p()
            capi.check(capi.lib().eskf_host_alloc(C.c_size_t(x32.nbytes), C.byref(px)))
            capi.check(capi.lib().eskf_host_alloc(C.c_size_t(t.nbytes), C.byref(pt)))
            C.memmove(px, x32.ctypes.data, x32.nbytes)
            C.memmove(pt, t.ctypes.data, t.nbytes)
            pinned.append((px, pt, len(t)))
        keep = pinned

        def feed(i):
            px, pt, n = pinned[i]
            od.feed_lidar_ptr(px.value, pt.value, n)

    state = {}
    dev_marks = []

    def start():
        ctx.sync()
        inf = od.info()
        state["dev0"] = inf.device_frame_ms_sum
        state["stage0"] = np.array(inf.stage_sum_ms)
        state["frames0"] = inf.frames
        state["launch0"] = od.launch_count()
        state["t0"] = time.perf_counter()

    walls, poses, iters = replay(od, scans, imu, feed, first_timed, start)
    ctx.sync()
    total_wall = time.perf_counter() - state["t0"]
    inf = od.info()
    n = int(inf.frames - state["frames0"])
    del dev_marks
    res = {"frames": n, "device_ms": inf.device_frame_ms_sum - state["dev0"],
           "stage_ms": ((np.array(inf.stage_sum_ms) - state["stage0"]) / max(n, 1)).tolist(),
           "wall_ms": float(np.sum(walls)), "walls": walls, "replay_wall_ms": 1e3 * total_wall,
           "launches": od.launch_count() - state["launch0"],
           "gn_iterations_mean": float(np.mean(iters)), "map_voxels": int(inf.map_voxels),
           "n_states": int(inf.n_states), "poses": poses,
           # the sweep is copied as float32 xyz (12 B / point); its per-point stamps (8 B / point) are read
           # by the HOST only (deskew segment table, ~40 segments x 104 B, which is what crosses the bus)
           "h2d": int(np.mean([len(t) * 12 for _, t in scans[first_timed:]])) + 40 * 104, "d2h": 16 * 8 + 64}
    od.close()
    for px, pt, _ in keep:
        capi.lib().eskf_host_free(px)
        capi.lib().eskf_host_free(pt)
    return res


def frame_leg(args, capi, odometry, D, rank, world, local_rank):
    """configs[1]: one replica per GPU (a 64k-point frame does not shard)."""
    first_timed = 1 + LEAD_IN + FRAME_WARMUP
    steps = FRAME_STEPS if world == 1 else min(FRAME_STEPS, 20)
    scans, imu = make_log(first_timed + steps, seed=43)
    D.barrier()
    res = gpu_sequence(capi, odometry, local_rank, scans, imu, first_timed, "resident")
    D.barrier()
    per_rank = D.gather(res["device_ms"] / max(res["frames"], 1))
    if world > 1:
        return {"note": "one independent replica of the configs[1] sequence per GPU (a frame does not shard); "
                        "device-timed ms/frame of every rank", "timed_frames": steps,
                "ms_per_frame_per_rank": per_rank, "config": frame_workload_config()} if rank == 0 else None
    e2e = gpu_sequence(capi, odometry, local_rank, scans, imu, first_timed, "e2e")
    host = gpu_sequence(capi, odometry, local_rank, scans, imu, first_timed, "host")
    dts, drs = zip(*[pose_delta(a, b) for a, b in zip(res["poses"], e2e["poses"])])
    w = np.array(e2e["walls"])
    out = {
        "metric": "ms_per_frame", "value": res["device_ms"] / steps, "unit": "ms",
        "timing": "value: CUDA events on the context's stream around the three stages of every timed "
                  "frame (sweeps resident in HBM), summed; e2e: wall clock around sweep delivery "
                  "(float32 H2D from pinned memory) + the three stages + pose read-back, per frame",
        "e2e": {"value": e2e["wall_ms"] / steps, "unit": "ms", "p50_ms": float(np.percentile(w, 50)),
                "p99_ms": float(np.percentile(w, 99)), "max_ms": float(w.max()),
                "frames_over_1ms": int((w > 1.0).sum()),
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "stage_ms": dict(zip(("preprocess", "filter_update", "map_update"), e2e["stage_ms"])),
                "replay_wall_ms_per_frame_incl_imu_propagation": e2e["replay_wall_ms"] / steps},
        "e2e_host_vector_classes": {
            "value": host["wall_ms"] / steps, "unit": "ms",
            "note": "the same frames through the drop-in classes with host std::vector clouds at "
                    "every call (Config::device_resident = false)"},
        "gpu_launches": res["launches"], "frames_per_s": 1e3 * steps / res["device_ms"],
        "stage_ms": dict(zip(("preprocess", "filter_update", "map_update"), res["stage_ms"])),
        "wall_ms_per_frame_resident": res["wall_ms"] / steps,
        "gn_iterations_mean": res["gn_iterations_mean"], "map_voxels": res["map_voxels"],
        "filter_states": res["n_states"],
        "resident_vs_e2e_pose_delta": {"max_m": max(dts), "max_rad": max(drs)},
        "config": frame_workload_config(),
    }
    if not args.no_cpu_baseline:
        n_cpu = first_timed + args.cpu_frames
        r = run_frame_oracle(scans[:n_cpu], imu, first_timed)
        od, oe = zip(*[pose_delta(a, b) for a, b in zip(r["poses"], res["poses"][:n_cpu])])
        out["cpu_baseline"] = {
            "value": r["ms_per_frame"], "unit": "ms", "cores": r["cores"], "kind": "port",
            "sample": f"{r['frames']} timed frames of the same log after the same {first_timed}-frame lead-in",
            "stage_ms": r["stage_ms"], "gn_iterations_mean": r["gn_iterations_mean"]}
        out["trajectory_match_vs_cpu"] = {"frames": n_cpu, "max_m": max(od), "max_rad": max(oe)}
        assert max(od) < 1e-5 and max(oe) < 1e-5, ("GPU trajectory off the CPU oracle's", max(od), max(oe))
    return out


def single_scan_leg(capi, local_rank, with_oracle):
    """configs[0] exactly as SURVEY.md 8(d) config 1: the full 64k-point scan downsampled at 0.5 m,
    aligned to a 20-scan local map from the 0.10 m / 1 deg perturbed pose."""
    rng = np.random.default_rng(42)
    scene = S.hall_scene()
    poses = S.arc_trajectory(21)
    T_il = S.default_T_il()
    ctx = capi.Context(local_rank)
    gmap = capi.Map(ctx, 0.5, 1000, 1 << 18)
    raw = capi.Cloud(ctx, 64000)
    ds = capi.Cloud(ctx, 64000)
    host = []
    for i in range(20):
        xyz, t = S.make_scan(scene, poses[i], rng)
        raw.upload_f32(xyz)
        raw.preprocess_into(ds, None, T_il, None, 0.5)
        if with_oracle:
            host.append(ds.download())
        gmap.insert_cloud(ds, poses[i])
    xyz, t = S.make_scan(scene, poses[20], rng)
    raw.upload_f32(xyz)
    raw.preprocess_into(ds, None, T_il, None, 0.5)
    guess = poses[20] @ S.perturbation()
    n_ds = ds.size()
    r = gmap.align_cloud(ds, guess, trace=True)
    for _ in range(3):
        gmap.align_cloud(ds, guess)
    ts = []
    for _ in range(10):
        ctx.sync()
        ctx.timer_start()
        gmap.align_cloud(ds, guess)
        ts.append(ctx.timer_stop())
    ms = float(np.median(ts))
    out = {"workload": "BASELINE.json configs[0]: one 64k-point scan, 0.5 m downsample, VGICP against a 20-scan "
                       "local map from a 0.10 m / 1 deg perturbed pose (SURVEY.md 8d config 1)",
           "n_downsampled": n_ds, "map_voxels": gmap.size(), "gn_iterations": r["iterations"],
           "converged": bool(r["converged"]), "align_ms": ms, "align_us_per_gn_iteration": 1e3 * ms / r["iterations"],
           "mpts_per_s_per_gn_iteration": n_ds * r["iterations"] / (ms * 1e-3) / 1e6}
    if with_oracle:
        import oracle as O
        om = O.Map(0.5, 1000)
        for (p, c, _), T in zip(host, poses[:20]):
            om.update(p, c, T, initialize=True)
        p, c, _ = ds.download()
        t0 = time.perf_counter()
        ro = om.align(p, c, guess)
        out["cpu_align_ms"] = 1e3 * (time.perf_counter() - t0)
        out["cpu_cores"] = O.num_threads()
        dt, dr = pose_delta(ro["T"], r["T"])
        out["vs_cpu_oracle"] = {"iterations_equal": ro["iterations"] == r["iterations"],
                                "ncorr_equal": bool(np.array_equal(ro["ncorr"], r["ncorr"])),
                                "pose_m": dt, "pose_rad": dr, "occupancy_equal": om.size() == gmap.size()}
        assert ro["iterations"] == r["iterations"] and dt < 1e-5 and dr < 1e-5
    del gmap, raw, ds
    ctx.close()
    return out


def batch_leg(capi, sharded, D, rank, world, local_rank, pairs=512, distinct_per_gpu=64, streams=8):
    """configs[4]: independent scan / local-map pairs, one shard of the batch per GPU, no collective;
    on each GPU one host thread keeps a registration in flight on each of `streams` contexts."""
    from scripts.batch_register import make_pair
    mine = list(sharded.shard_batch(pairs, rank, world))
    ctxs = [capi.Context(local_rank) for _ in range(streams)]
    jobs = []
    for n, k in enumerate(mine[:distinct_per_gpu]):
        mp, mc, sp, sc, guess = make_pair(k, 300_000, 15_000)
        c = ctxs[n % streams]
        gm = capi.Map(c, 0.5, 1000, 1 << 15)
        gm.insert(mp, mc, np.eye(4))
        jobs.append((gm, capi.Cloud(c, len(sp)).upload(sp, sc), guess))
    for c in ctxs:
        c.sync()
    order = [jobs[i % len(jobs)] for i in range(len(mine))]
    args3 = ([j[0] for j in order], [j[1] for j in order], [j[2] for j in order])
    capi.align_batch(ctxs, *[a[:streams] for a in args3])  # warm-up
    D.barrier()
    t0 = time.perf_counter()
    rs = capi.align_batch(ctxs, *args3)
    dt = time.perf_counter() - t0
    dt = D.max(dt)[0]
    ok = all(r["converged"] for r in rs)
    del jobs, order, args3
    for c in ctxs:
        c.close()
    return {"workload": f"BASELINE.json configs[4]: {pairs} scan / local-map registrations (15k-point scan vs 300k-point "
                        f"map at 0.5 m, run to convergence), {len(mine)} per GPU over {min(distinct_per_gpu, len(mine))} "
                        "distinct pairs, no collective", "registrations_per_s": pairs / dt, "seconds": dt,
            "streams_per_gpu": streams, "all_converged": bool(ok),
            "gn_iterations_mean": float(np.mean([r["iterations"] for r in rs]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame", action="store_true", help="skip the configs[1] / configs[0] sub-records")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed registrations of the cpu_baseline leg")
    ap.add_argument("--cpu-frames", type=int, default=5, help="timed frames of the frame sub-record's CPU leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        impl_reference(args, rank, world)
        return

    from eskf_lio_b200 import capi, odometry, sharded
    D = Dist(world, local_rank)
    local_rank = D.local
    D.barrier()
    d = dense_leg(args, capi, sharded, D, rank, world, local_rank)
    D.barrier()
    total_pts_iters = DENSE_SRC * DENSE_ITERS * args.steps
    value = total_pts_iters / (d["ms_dev"] * 1e-3) / 1e6
    e2e_value = total_pts_iters / (d["ms_e2e"] * 1e-3) / 1e6
    peak, peak_src, _ = peaks()
    line = None
    if rank == 0:
        ms_it = d["ms_launch"] / DENSE_ITERS
        bytes_launch = d["shard_points"] * ALG_BYTES_PER_POINT_ITER * DENSE_ITERS
        achieved = bytes_launch / (d["ms_launch"] * 1e-3) / 1e9
        hit_bytes = sum(48 * DENSE_SRC + 88 * h for h in d["ncorr_fixed"]) / world
        traffic, traffic_src = dense_traffic()
        line = {
            "metric": "mpts_per_s_per_gn_iteration", "value": value, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": d["ms_dev"] / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 positions/keys + f32 per-point algebra + f64 accumulation",
            "data": "synthetic", "config": dense_workload_config(world),
            "timing": "value: CUDA events on the context's stream around the K registrations (shards resident in "
                      "HBM), max over ranks; e2e: wall clock around K x {upload the shard from pinned host "
                      "memory, register, read the pose back}, max over ranks; barrier + synchronize on both sides",
            "ms_per_gn_iteration": d["ms_dev"] / args.steps / DENSE_ITERS,
            "e2e": {"value": e2e_value, "unit": "Mpts/s", "h2d_bytes_per_step": d["h2d"],
                    "d2h_bytes_per_step": d["d2h"], "ms_per_step": d["ms_e2e"] / args.steps,
                    "note": "bytes are per rank: the point range of this rank as fp64 xyz (24 B) + row-major "
                            "3x3 covariance (72 B) per point"},
            "gpu_launches": d["launches"], "clocks": d["clocks"], "map_voxels": d["n_vox"],
            "parity": d["parity"],
            "roofline": {
                "bound": "hbm", "kernel": "align_kernel<float> (fused transform + voxel lookup + "
                                          "J^T W J / J^T W r + reduction + on-device solve + NVLink exchange)",
                "kernel_shape": d["kernel_shape"] + " (chosen by the context's one-off timing of both on this device)",
                "workload": f"{d['shard_points']} source points per GPU vs the {d['n_vox']}-voxel map, "
                            f"{DENSE_ITERS} GN iterations per launch",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "frac_of_8TBs_spec": achieved / 8000.0,
                "traffic": traffic if world == 1 else None, "traffic_source": traffic_src if world == 1 else None,
                "algorithmic_bytes_per_launch": bytes_launch,
                "hit_weighted": {"note": "only the points that find a voxel need the 64 B record and the 24 B "
                                         "source covariance; bytes = 48 N + 88 hits per iteration (per GPU)",
                                 "bytes_per_launch": hit_bytes,
                                 "frac": hit_bytes / (d["ms_launch"] * 1e-3) / 1e9 / peak,
                                 "hit_rate": float(np.mean(d["ncorr_fixed"])) / DENSE_SRC},
                "ms_per_launch": d["ms_launch"], "ms_per_gn_iteration": ms_it,
                "mpts_per_s_per_gn_iteration": DENSE_SRC / (ms_it * 1e-3) / 1e6,
                "timing": "CUDA events on the launching stream (eskf_ctx_timer_*), median of 5 single "
                          "launches, max over ranks"},
        }
        if world > 1:
            line["strong_scaling_in_this_run"] = {
                "unsharded_ms_per_gn_iteration_rank0": d["unsharded_ms_launch_rank0"] / DENSE_ITERS,
                "sharded_ms_per_gn_iteration": ms_it,
                "note": "both measured as single launches in this process group"}
            if "nccl_ms_launch" in d:
                line["nccl_baseline"] = {"ms_per_gn_iteration": d["nccl_ms_launch"] / DENSE_ITERS,
                                         "note": "eskf_align_cloud_sharded: two launches + an NCCL all-reduce of "
                                                 "the 28 sums per iteration"}
            if "weak_ms_launch" in d:
                wit = d["weak_ms_launch"] / DENSE_ITERS
                line["weak"] = {"points_per_gpu": DENSE_SRC, "ms_per_gn_iteration": wit,
                                "mpts_per_s_per_gn_iteration": DENSE_SRC * world / (wit * 1e-3) / 1e6}
    D.barrier()
    if not args.no_batch:
        b = batch_leg(capi, sharded, D, rank, world, local_rank)
        if rank == 0:
            line["batch"] = b
    D.barrier()
    if not args.no_frame:
        f = frame_leg(args, capi, odometry, D, rank, world, local_rank)
        if rank == 0:
            line["frame"] = f
            if world == 1:
                line["single_scan"] = single_scan_leg(capi, local_rank, not args.no_cpu_baseline)
    D.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        orc = DenseOracle()
        r = orc.run(args.cpu_steps, 1)
        line["cpu_baseline"] = {"value": r["mpts_per_s"], "unit": "Mpts/s", "cores": r["cores"], "kind": "port",
                                "sample": cpu_sample_text(r), "map_build_s": r["map_build_s"]}
        # parity of the GPU path against the oracle on that sample (run to convergence)
        m = min(CPU_SAMPLE, DENSE_SRC)
        p, c = d["full0"]
        ro = orc.map.align(p[:m], c[:m], orc.guess)
        rg = d["sample_gpu"]
        dt, dr = pose_delta(ro["T"], rg["T"])
        line["parity"]["vs_cpu_oracle_on_sample"] = {
            "sample_points": m, "iterations_equal": ro["iterations"] == rg["iterations"],
            "ncorr_equal": bool(np.array_equal(ro["ncorr"], rg["ncorr"])),
            "occupancy_equal": orc.map.size() == d["n_vox"], "pose_m": dt, "pose_rad": dr}
        assert ro["iterations"] == rg["iterations"] and dt < 1e-5 and dr < 1e-5, "GPU vs oracle parity"
        assert np.array_equal(ro["ncorr"], rg["ncorr"]), "GPU vs oracle correspondence counts"
    D.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    D.close()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE align_kernel launch of this very workload, read from
# the committed summary of the `ncu --set full` capture (profiles/); None when there is no capture
DENSE_TRAFFIC_FILE = "profiles/r2_prof_align_ncu.md"


def dense_traffic():
    path = os.path.join(ROOT, DENSE_TRAFFIC_FILE)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        with open(path) as f:
            for ln in f:
                cells = [c.strip() for c in ln.strip().strip("|").split("|")]
                if len(cells) == 3 and cells[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(cells[2].replace(",", "")) * scale[cells[1]]
                    seen += 1
    except (OSError, KeyError, ValueError):
        return None, None
    if seen != 2:
        return None, None
    return int(total), DENSE_TRAFFIC_FILE + " (ncu --set full, one launch)"


if __name__ == "__main__":
    main()
