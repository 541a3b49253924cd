"""N>1 host logic on CPU (gloo, world_size 2): point-range sharding + the
per-iteration all-reduce of the 28 sums reproduce the unsharded linearisation,
and every rank ends up with the identical step.  (The compute here is the CPU
oracle standing in for the kernel; the GPU path is covered by -m gpu tests.)"""
import os

import numpy as np
import pytest

from eskf_lio_b200.sharded import shard_batch, shard_range


def test_shard_range_tiles_exactly():
    for n in (0, 1, 7, 64000, 2_000_001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    assert list(shard_batch(10, 1, 4)) == [3, 4, 5]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from eskf_lio_b200 import synth as S
    O.set_num_threads(1)
    rng = np.random.default_rng(5)
    scene = S.hall_scene()
    poses = S.arc_trajectory(3)
    T_il = S.default_T_il()
    om = O.Map(0.5, 1000)
    clouds = []
    for T in poses:
        xyz, t = S.make_scan(scene, T, rng)
        clouds.append(O.preprocess(xyz[::8], t[::8], T_il, None, 0.5))
    for (p, c, _), T in zip(clouds[:2], poses[:2]):
        om.update(p, c, T, initialize=True)        # map replicated on every rank
    p, c, _ = clouds[2]
    pg, cg = O.transform_cloud(p, c, poses[2] @ S.perturbation())
    b, e = shard_range(len(pg), rank, world)
    H, bb, hit, nc = om.linearize(pg[b:e], cg[b:e])
    buf = torch.from_numpy(np.concatenate([H.ravel(), bb, [float(nc)]]))
    dist.all_reduce(buf)                            # the per-iteration exchange
    Hs, bs, ncs = buf[:36].numpy().reshape(6, 6), buf[36:42].numpy(), int(buf[42])
    step = O.se3_to_SE3(O.ldlt_solve6(Hs, -bs))
    Hf, bf, _, ncf = om.linearize(pg, cg)
    q.put((rank, float(np.abs(Hs - Hf).max() / np.abs(Hf).max()), float(np.abs(bs - bf).max()),
           ncs == ncf, step.tobytes()))
    dist.destroy_process_group()


def test_sharded_linearisation_gloo_world2(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dH, db, same_nc, _ in res:
        assert dH < 1e-12 and db < 1e-6 and same_nc
    assert res[0][4] == res[1][4]  # identical step on every rank, bit for bit
