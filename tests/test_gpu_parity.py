"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through
the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): voxel keys, occupancy and correspondence sets
bit-exact; per-iteration H/b within 1e-4 relative (norm-wise); final pose
within 1e-5 rad and 1e-5 m."""
import numpy as np
import pytest

from eskf_lio_b200 import capi, synth as S
from gpu_common import Frames, pose_err, rel_err

pytestmark = pytest.mark.gpu

H_TOL = 1e-4      # per-iteration H / b, norm-wise relative
POSE_T_TOL = 1e-5  # metres
POSE_R_TOL = 1e-5  # radians


@pytest.fixture(scope="module")
def ctx():
    return capi.Context(0)


@pytest.fixture(scope="module")
def frames(oracle):
    return Frames(oracle, n_scans=5, seed=11, decim=2, voxel=0.5)


def b_rel(bg, bo, Ho):
    # b -> 0 at convergence: relative to max(|b|, |H| * 1e-3 m)
    den = max(np.linalg.norm(bo), np.linalg.norm(Ho) * 1e-3 * 1e-3)
    return float(np.linalg.norm(bg - bo) / den)


# ------------------------------------------------------------------ clouds
def test_cloud_roundtrip_and_transform_bit_exact(ctx, oracle, frames):
    p, c = frames.ds[0]
    cl = capi.Cloud(ctx).upload(p, c)
    x, cv, _ = cl.download()
    np.testing.assert_array_equal(x, p)
    np.testing.assert_array_equal(cv, c)
    T = frames.poses[3] @ S.perturbation()
    cl.transform(T)
    x, cv, _ = cl.download()
    ox, oc = oracle.transform_cloud(p, c, T)
    np.testing.assert_array_equal(x, ox)     # fp64, same evaluation order, no FMA
    np.testing.assert_array_equal(cv, oc)


def test_upload_f32_matches_widening(ctx, frames):
    xyz = frames.raw[0][0][:1000]
    cl = capi.Cloud(ctx).upload_f32(xyz.astype(np.float32))
    x, _, _ = cl.download(want_cov=False)
    np.testing.assert_array_equal(x, xyz)


# --------------------------------------------------------------------- map
@pytest.mark.parametrize("voxel", [0.1, 0.3, 0.5, 1.0])
def test_voxel_keys_bit_exact(ctx, oracle, frames, voxel):
    p, c = frames.ds[1]
    T = frames.poses[1]
    gm = capi.Map(ctx, voxel, 1000, 1 << 12)
    gm.insert(p, c, T)
    pw, _ = oracle.transform_cloud(p, c, T)
    # add adversarial points on voxel faces
    edge = np.array([[-0.1, -0.5, -0.0], [0.5, 0.4999999999999999, 0.0], [0.3, 0.6, 0.9],
                     [voxel, -voxel, 2 * voxel], [np.nextafter(voxel, 0), voxel * 3, -voxel * 3]])
    q = np.vstack([pw, edge])
    keys, hit, count, mean, cov = gm.query(q)
    np.testing.assert_array_equal(keys, oracle.voxel_index(q, voxel))
    assert hit[:len(pw)].all()


def test_map_insert_matches_oracle(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    ok, oc, omean, ocov = om.export()
    gk, gc, gmean, gcov = gm.export()
    assert gm.size() == om.size()
    np.testing.assert_array_equal(gk, ok)                       # occupancy bit-exact
    np.testing.assert_array_equal(gc.astype(np.uint64), oc)     # counts bit-exact
    # running mean / covariance folded in input order with the reference's expression
    np.testing.assert_array_equal(gmean, omean)
    np.testing.assert_array_equal(gcov, ocov)


def test_map_cap_and_long_runs(ctx, oracle):
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1.0, 1.0, size=(6000, 3))
    A = rng.normal(size=(6000, 3, 3))
    cov = A @ A.transpose(0, 2, 1) + np.eye(3)
    om = oracle.Map(1.0, 7)
    gm = capi.Map(ctx, 1.0, 7, 1 << 10)
    for k in range(3):
        sl = slice(2000 * k, 2000 * (k + 1))
        om.insert(pts[sl], cov[sl])
        gm.insert(pts[sl], cov[sl], np.eye(4))
    ok, oc, omean, ocov = om.export()
    gk, gc, gmean, gcov = gm.export()
    np.testing.assert_array_equal(gk, ok)
    np.testing.assert_array_equal(gc.astype(np.uint64), oc)
    assert gc.max() == 7
    np.testing.assert_array_equal(gmean, omean)
    np.testing.assert_array_equal(gcov, ocov)


def test_map_growth_from_tiny_table(ctx, oracle, frames):
    om = oracle.Map(0.1, 1000)
    gm = capi.Map(ctx, 0.1, 1000, 16)
    cap0 = gm.capacity()
    for (p, c), T in zip(frames.ds[:3], frames.poses[:3]):
        om.update(p, c, T, initialize=True)
        gm.insert(p, c, T)
    assert gm.capacity() > cap0
    assert gm.size() == om.size()
    gk, gc, gmean, _ = gm.export()
    ok, oc, omean, _ = om.export()
    np.testing.assert_array_equal(gk, ok)
    np.testing.assert_array_equal(gc.astype(np.uint64), oc)
    np.testing.assert_array_equal(gmean, omean)


def test_map_evict_matches_oracle(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 3)
    pos = frames.poses[2][:3, 3]
    for thr in (25.0, 12.0):
        assert gm.evict(pos, thr) == om.evict(pos, thr)
        gk, gc, gmean, gcov = gm.export()
        ok, oc, omean, ocov = om.export()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc.astype(np.uint64), oc)
        np.testing.assert_array_equal(gmean, omean)
    # strict '>' at exactly the threshold
    om2 = oracle.Map(1.0, 10)
    gm2 = capi.Map(ctx, 1.0, 10, 64)
    pts = np.array([[0.5, 0.5, 0.5], [3.5, 0.5, 0.5], [4.5, 0.5, 0.5]])
    cv = np.repeat(np.eye(3)[None], 3, axis=0)
    om2.insert(pts, cv)
    gm2.insert(pts, cv, np.eye(4))
    assert gm2.evict([0.5, 0.5, 0.5], 3.0) == om2.evict([0.5, 0.5, 0.5], 3.0) == 1
    np.testing.assert_array_equal(gm2.export()[0], [[0, 0, 0], [3, 0, 0]])
    # inserting after an eviction still works
    gm2.insert(pts, cv, np.eye(4))
    assert gm2.size() == 3


def test_insert_list_path_equals_sorted_path(ctx, oracle, frames):
    """Per-frame batches take the sort-free insert (pending lists folded in point-index order), large
    ones the radix-sort path: both must give bit-identical maps — and equal the oracle."""
    rng = np.random.default_rng(8)
    # a batch with long per-voxel lists (> 32 points of one batch in a voxel) and a cap that bites
    pts = np.concatenate([rng.uniform(0.0, 1.0, size=(300, 3)), rng.uniform(-4.0, 4.0, size=(3000, 3))])
    A = rng.normal(size=(len(pts), 3, 3))
    cov = A @ A.transpose(0, 2, 1) + np.eye(3)
    maps = []
    for sorted_path in (0, 1):
        ctx.set_option("map_insert_sorted", sorted_path)
        try:
            gm = capi.Map(ctx, 1.0, 120, 1 << 10)
            for (p, c), T in zip(frames.ds[:3], frames.poses[:3]):
                gm.insert(p, c, T)
            gm.insert(pts, cov, frames.poses[1])
            gm.insert(pts[::-1].copy(), cov[::-1].copy(), np.eye(4))
            maps.append(gm.export())
        finally:
            ctx.set_option("map_insert_sorted", 0)
    for u, v in zip(*maps):
        np.testing.assert_array_equal(u, v)
    om = oracle.Map(1.0, 120)
    for (p, c), T in zip(frames.ds[:3], frames.poses[:3]):
        om.update(p, c, T, initialize=True)
    om.update(pts, cov, frames.poses[1], initialize=True)
    om.update(pts[::-1].copy(), cov[::-1].copy(), np.eye(4), initialize=True)
    ok, oc, omean, ocov = om.export()
    np.testing.assert_array_equal(maps[0][0], ok)
    np.testing.assert_array_equal(maps[0][1].astype(np.uint64), oc)
    np.testing.assert_array_equal(maps[0][2], omean)
    np.testing.assert_array_equal(maps[0][3], ocov)
    assert oc.max() == 120


@pytest.mark.parametrize("sorted_path", [0, 1])
def test_key_range_error(ctx, sorted_path):
    """A point whose voxel coordinate leaves the 21-bit key range is dropped and reported on BOTH insert
    paths (the sorted path used to fold it into an aliased voxel), the in-range points of the
    batch land where they belong."""
    ctx.set_option("map_insert_sorted", sorted_path)
    try:
        gm = capi.Map(ctx, 0.01, 10, 64)
        pts = np.array([[0.0, 0.0, 0.0], [2.0e4, 0.0, 0.0], [0.015, 0.0, 0.0]])  # 2e6 voxels > 2^20
        gm.insert(pts, np.repeat(np.eye(3)[None], 3, axis=0), np.eye(4))
        with pytest.raises(capi.EskfError) as e:
            gm.size()
        assert e.value.status == 5
        keys, hit, count, mean, _ = gm.query(pts[[0, 2]])
        assert hit.all() and count.tolist() == [1, 1]
        np.testing.assert_array_equal(mean, pts[[0, 2]])
        assert keys.tolist() == [[0, 0, 0], [1, 0, 0]]
    finally:
        ctx.set_option("map_insert_sorted", 0)


# -------------------------------------------------------------- preprocess
@pytest.mark.parametrize("cap", [1, 8])
def test_voxelize_one_cluster_equals_grid_wide_kernel(ctx, oracle, frames, cap):
    """A sweep is voxelised + radix-sorted by ONE thread-block cluster (hardware barriers; option
    "vox_cluster", 16 CTAs or capped at 8) — every output must be bit-identical with the grid-wide kernel's
    (software barriers), for the preprocessor (full sweep, a 2.5k-point one, deskew + crop) and for the
    sorted map-insert path, and equal to the oracle's."""
    xyz, t = frames.raw[3]
    states = _deskew_states(t)
    r = np.linalg.norm(xyz, axis=1)
    cases = [(xyz, t, states, 0.3, (0.0, 0.0)), (xyz[:2500], t[:2500], None, 0.5, (0.0, 0.0)),
             (xyz, t, states, 0.5, (float(np.quantile(r, 0.1)), float(np.quantile(r, 0.9)))), (xyz[:1], t[:1], None, 0.5, (0.0, 0.0)),
             (xyz[:6000], t[:6000], None, 0.0007, (0.0, 0.0)), (xyz[:30000], t[:30000], states, 0.3, (0.0, 0.0))]   # > 16 bits per axis: key and index in separate arrays
    out = {}
    try:
        for mode in (0, cap):
            ctx.set_option("vox_cluster", mode)
            res = []
            for x, tt, st, v, (mn, mx) in cases:
                ctx.set_range_crop(mn, mx)
                res.append(ctx.preprocess(x, tt if st is not None else None, frames.T_il, st, v))
            ctx.set_range_crop(0.0, 0.0)
            ctx.set_option("map_insert_sorted", 1)
            gm = capi.Map(ctx, 0.5, 50, 1 << 10)
            for (p, c), T in zip(frames.ds[:3], frames.poses[:3]):
                gm.insert(p, c, T)
            ctx.set_option("map_insert_sorted", 0)
            out[mode] = (res, gm.export())
    finally:
        ctx.set_option("vox_cluster", 1)
        ctx.set_option("map_insert_sorted", 0)
        ctx.set_range_crop(0.0, 0.0)
    for (p0, c0, s0), (p1, c1, s1) in zip(out[0][0], out[cap][0]):
        np.testing.assert_array_equal(s0, s1)
        np.testing.assert_array_equal(p0, p1)
        np.testing.assert_array_equal(c0, c1)
    for u, v in zip(out[0][1], out[cap][1]):
        np.testing.assert_array_equal(u, v)
    x, tt, st, v, _ = cases[0]
    op, oc, osrc = oracle.preprocess(x, tt, frames.T_il, st, v)
    np.testing.assert_array_equal(out[cap][0][0][2], osrc)
    np.testing.assert_array_equal(out[cap][0][0][0], op)
    assert np.abs(out[cap][0][0][1] - oc).max() < 1e-7


@pytest.mark.parametrize("voxel", [0.5, 0.3])
def test_downsample_and_covariances(ctx, oracle, frames, voxel):
    xyz, t = frames.raw[2]
    op, oc, osrc = oracle.preprocess(xyz, t, frames.T_il, None, voxel)
    gp, gc, gsrc = ctx.preprocess(xyz, t, frames.T_il, None, voxel)
    np.testing.assert_array_equal(gsrc, osrc)     # kept set (first point per voxel) bit-exact
    np.testing.assert_array_equal(gp, op)         # transformed positions bit-exact
    err = np.abs(gc - oc).reshape(len(gc), -1).max(axis=1)
    assert np.median(err) < 1e-12
    assert err.max() < 1e-7, f"worst covariance mismatch {err.max()} at {err.argmax()}"


def test_downsample_small_and_degenerate_inputs(ctx, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 29, 31, 200):
        pts = rng.normal(size=(n, 3)) * 2.0
        op, oc, osrc = oracle.downsample_cov(pts, 0.5)
        gp, gc, gsrc = ctx.downsample_cov(pts, 0.5)
        np.testing.assert_array_equal(gsrc, osrc)
        np.testing.assert_array_equal(gp, op)
        np.testing.assert_allclose(gc, oc, atol=1e-7)
    # all points in one voxel
    pts = rng.uniform(0.01, 0.49, size=(500, 3))
    gp, gc, gsrc = ctx.downsample_cov(pts, 0.5)
    assert gsrc.tolist() == [0]
    op, oc, _ = oracle.downsample_cov(pts, 0.5)
    np.testing.assert_allclose(gc, oc, atol=1e-9)


def test_knn_search_corner_paths(ctx, oracle):
    """The k-NN search beyond its common path: (1) far-spread sparse points (levels above the
    block-range tables: binary-search fallback, up to the whole-cloud level), (2) the serial-insertion
    overflow path of the selection, (3) a dense blob next to an isolated point."""
    rng = np.random.default_rng(12)
    # (1) 400 points over 600 m with 0.1 m voxels: every neighbourhood needs coarse levels
    pts = rng.uniform(-300.0, 300.0, size=(400, 3))
    op, oc, osrc = oracle.downsample_cov(pts, 0.1)
    gp, gc, gsrc = ctx.downsample_cov(pts, 0.1)
    np.testing.assert_array_equal(gsrc, osrc)
    np.testing.assert_allclose(gc, oc, atol=1e-7)
    # (2) forced overflow: with room for only 4 candidates every query takes the serial-insertion
    #     path; same exact neighbour sets in the same order => bit-identical covariances
    pts = rng.normal(size=(4000, 3)) * np.array([3.0, 3.0, 0.2])
    ref = ctx.downsample_cov(pts, 0.5)
    ctx.set_option("knn_buffer", 4)
    try:
        alt = ctx.downsample_cov(pts, 0.5)
    finally:
        ctx.set_option("knn_buffer", 128)
    for u, v in zip(ref, alt):
        np.testing.assert_array_equal(u, v)
    np.testing.assert_allclose(ref[1], oracle.downsample_cov(pts, 0.5)[1], atol=1e-7)
    with pytest.raises(capi.EskfError):
        ctx.set_option("no_such_option", 1)
    # (3) blob + loner
    pts = np.concatenate([rng.normal(size=(5000, 3)) * 0.05, [[40.0, -35.0, 3.0]]])
    op, oc, osrc = oracle.downsample_cov(pts, 0.3)
    gp, gc, gsrc = ctx.downsample_cov(pts, 0.3)
    np.testing.assert_array_equal(gsrc, osrc)
    np.testing.assert_allclose(gc, oc, atol=1e-7)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_knn_ties_on_a_lattice(ctx, oracle, seed):
    """Exact lattices make whole shells of neighbours equidistant; the 30th neighbour falls inside a
    shell, so the (distance, index) tie order decides the set — and with it the covariance."""
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(14), np.arange(11), np.arange(5), indexing="ij"), -1).reshape(-1, 3)
    pts = (g * 0.125)[rng.permutation(len(g))] + rng.integers(-3, 4, size=3) * 0.5
    for voxel in (0.5, 0.3):
        op, oc, osrc = oracle.downsample_cov(pts, voxel)
        gp, gc, gsrc = ctx.downsample_cov(pts, voxel)
        np.testing.assert_array_equal(gsrc, osrc)
        np.testing.assert_array_equal(gp, op)
        np.testing.assert_allclose(gc, oc, atol=1e-9)


def test_preprocess_with_deskew(ctx, oracle, frames):
    xyz, t = frames.raw[1]
    t0, t1 = t[0], t[-1]
    ts = t0 - 0.006 + 0.0025 * np.arange(48)
    assert ts[-1] > t1
    s = ts - t0
    pos = np.stack([1.2 * s, 0.3 * s * s, 0.05 * np.sin(8 * s)], axis=1)
    ang = 0.4 * s
    quat = np.stack([0.02 * np.sin(ang), 0.01 * np.sin(ang), np.sin(ang / 2), np.cos(ang / 2)], axis=1)
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    # long history before the sweep, like the never-trimmed states_ deque
    hist_ts = t0 - 0.006 - 0.0025 * np.arange(400, 0, -1)
    ts = np.concatenate([hist_ts, ts])
    pos = np.concatenate([np.zeros((400, 3)), pos])
    quat = np.concatenate([np.tile([0, 0, 0, 1.0], (400, 1)), quat])
    states = (ts, pos, quat)
    op, oc, osrc = oracle.preprocess(xyz, t, frames.T_il, states, 0.5)
    gp, gc, gsrc = ctx.preprocess(xyz, t, frames.T_il, states, 0.5)
    np.testing.assert_array_equal(gsrc, osrc)
    np.testing.assert_array_equal(gp, op)
    assert np.abs(gc - oc).max() < 1e-7


def _deskew_states(t):
    t0, t1 = t[0], max(t[-1], t.max())
    ts = t0 - 0.006 + 0.0025 * np.arange(int((t1 - t0 + 0.02) / 0.0025) + 4)
    s = ts - t0
    pos = np.stack([1.2 * s, 0.3 * s * s, 0.05 * np.sin(8 * s)], axis=1)
    ang = 0.4 * s
    quat = np.stack([0.02 * np.sin(ang), 0.01 * np.sin(ang), np.sin(ang / 2), np.cos(ang / 2)], axis=1)
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    return ts, pos, quat


def test_preprocess_with_unsorted_stamps(ctx, oracle, frames):
    """stamps that are not non-decreasing (ring-major / merged sweeps): the segments must be
    the reference's forward linear scan (src/CloudPreprocessor.cpp:54-61), which the oracle restates and
    tests/test_reference_shim.py holds to the reference's own sources."""
    xyz, t = frames.raw[1]
    rng = np.random.default_rng(3)
    blocks = np.array_split(np.arange(len(t)), 37)
    tu = t[np.concatenate([blocks[k] for k in rng.permutation(len(blocks))])].copy()
    assert np.any(np.diff(tu) < 0)
    states = _deskew_states(t)
    op, oc, osrc = oracle.preprocess(xyz, tu, frames.T_il, states, 0.5)
    gp, gc, gsrc = ctx.preprocess(xyz, tu, frames.T_il, states, 0.5)
    np.testing.assert_array_equal(gsrc, osrc)
    np.testing.assert_array_equal(gp, op)
    assert np.abs(gc - oc).max() < 1e-7
    sp, _, ssrc = oracle.preprocess(xyz, t, frames.T_il, states, 0.5)
    assert len(ssrc) != len(osrc) or not np.array_equal(sp, op)   # (a different segmentation than the sorted one)
    # the caller's answer instead of the per-call check (option "stamps_sorted"): same results
    assert not capi.stamps_sorted(tu) and capi.stamps_sorted(t)
    try:
        ctx.set_option("stamps_sorted", 0)
        hp, _, hsrc = ctx.preprocess(xyz, tu, frames.T_il, states, 0.5)
        np.testing.assert_array_equal(hsrc, osrc)
        np.testing.assert_array_equal(hp, op)
        ctx.set_option("stamps_sorted", 1)
        hp, _, hsrc = ctx.preprocess(xyz, t, frames.T_il, states, 0.5)
        np.testing.assert_array_equal(hsrc, ssrc)
        np.testing.assert_array_equal(hp, sp)
        with pytest.raises(capi.EskfError):
            ctx.set_option("stamps_sorted", 2)
    finally:
        ctx.set_option("stamps_sorted", -1)


@pytest.mark.parametrize("with_states", [False, True])
def test_range_crop_matches_oracle(ctx, oracle, frames, with_states):
    """north_star's range crop (defined by the oracle, off by default): kept set, source indices and
    positions bit-exact, covariances within 1e-7; edge cases: everything cropped, one-sided bounds."""
    xyz, t = frames.raw[2]
    states = _deskew_states(t) if with_states else None
    r = np.linalg.norm(xyz, axis=1)
    lo, hi = float(np.quantile(r, 0.15)), float(np.quantile(r, 0.85))
    try:
        for mn, mx in ((lo, hi), (lo, 0.0), (0.0, hi), (float(r[17]), float(r[17]) * (1 + 1e-15) + 20.0)):
            ctx.set_range_crop(mn, mx)
            op, oc, osrc = oracle.preprocess(xyz, t, frames.T_il, states, 0.5, min_range=mn, max_range=mx)
            gp, gc, gsrc = ctx.preprocess(xyz, t, frames.T_il, states, 0.5)
            assert 0 < len(osrc) < len(oracle.preprocess(xyz, t, frames.T_il, states, 0.5)[2])
            np.testing.assert_array_equal(gsrc, osrc)          # indices into the uncropped sweep
            np.testing.assert_array_equal(gp, op)
            assert np.abs(gc - oc).max() < 1e-7
            assert np.all(r[gsrc] >= mn * (1 - 1e-12)) and (mx == 0.0 or np.all(r[gsrc] <= mx * (1 + 1e-12)))
        ctx.set_range_crop(1e6, 0.0)                             # nothing survives
        gp, gc, gsrc = ctx.preprocess(xyz, t, frames.T_il, states, 0.5)
        assert len(gp) == 0 and len(gsrc) == 0
        with pytest.raises(capi.EskfError):
            ctx.set_range_crop(5.0, 1.0)
        ctx.set_range_crop(0.0, 0.0)                             # off again: the plain result
        op, oc, osrc = oracle.preprocess(xyz, t, frames.T_il, states, 0.5)
        gp, gc, gsrc = ctx.preprocess(xyz, t, frames.T_il, states, 0.5)
        np.testing.assert_array_equal(gsrc, osrc)
        np.testing.assert_array_equal(gp, op)
    finally:
        ctx.set_range_crop(0.0, 0.0)


# ------------------------------------------------------------ registration
def test_linearize_correspondences_and_Hb(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    pg, cg = oracle.transform_cloud(p, c, guess)
    Ho, bo, hito, nco = om.linearize(pg, cg)
    # identity pose on pre-transformed points: positions reach the kernel bit-exact
    Hg, bg, hitg, ncg = ctx.linearize(gm, pg, cg)
    np.testing.assert_array_equal(hitg, hito)      # correspondence set bit-exact
    assert ncg == nco and nco > 1000
    assert rel_err(Hg, Ho) < H_TOL and b_rel(bg, bo, Ho) < H_TOL
    # fp64 per-point math variant: limited only by the fp32 table
    Hd, bd, hitd, _ = ctx.linearize(gm, pg, cg, fp64_math=True)
    np.testing.assert_array_equal(hitd, hito)
    assert rel_err(Hd, Ho) < 1e-5 and b_rel(bd, bo, Ho) < 1e-5
    # pose applied on the device: same keys as the oracle's Transform
    Hg2, bg2, hitg2, _ = ctx.linearize(gm, p, c, T=guess)
    np.testing.assert_array_equal(hitg2, hito)
    assert rel_err(Hg2, Ho) < H_TOL and b_rel(bg2, bo, Ho) < H_TOL


def test_align_pose_iterations_and_trace(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    ro = om.align(p, c, guess)
    rg = ctx.align(gm, p, c, guess)
    assert ro["converged"] and rg["converged"]
    assert rg["iterations"] == ro["iterations"]
    dt, dr = pose_err(ro["T"], rg["T"])
    assert dt < POSE_T_TOL and dr < POSE_R_TOL, (dt, dr)
    np.testing.assert_array_equal(rg["ncorr"], ro["ncorr"])
    for k in range(ro["iterations"]):
        assert rel_err(rg["H"][k], ro["H"][k]) < H_TOL, k
        assert b_rel(rg["b"][k], ro["b"][k], ro["H"][k]) < H_TOL, k
    gt_t, gt_r = pose_err(frames.poses[4], rg["T"])
    assert gt_t < 0.02 and gt_r < 0.01


def test_align_teacher_forced_correspondences_every_iteration(ctx, oracle, frames):
    """Feed the oracle's own per-iteration clouds to the kernel: the
    correspondence set of EVERY Gauss-Newton iteration must match bit for bit."""
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation(dt=(0.2, 0.1, -0.05), angle_deg=2.0)
    ro = om.align(p, c, guess)
    pk, ck = oracle.transform_cloud(p, c, guess)
    for k in range(ro["iterations"]):
        Ho, bo, hito, nco = om.linearize(pk, ck)
        Hg, bg, hitg, ncg = ctx.linearize(gm, pk, ck)
        np.testing.assert_array_equal(hitg, hito)
        assert ncg == nco == ro["ncorr"][k]
        assert rel_err(Hg, Ho) < H_TOL and b_rel(bg, bo, Ho) < H_TOL
        pk, ck = oracle.transform_cloud(pk, ck, ro["step"][k])


def test_align_direct7_extension(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    pg, cg = oracle.transform_cloud(p, c, guess)
    Ho, bo, hito, nco = om.linearize(pg, cg, 7)
    Hg, bg, hitg, ncg = ctx.linearize(gm, pg, cg, neighbor_mode=7)
    np.testing.assert_array_equal(hitg, hito)
    assert ncg == nco
    assert rel_err(Hg, Ho) < H_TOL and b_rel(bg, bo, Ho) < H_TOL
    ro = om.align(p, c, guess, neighbor_mode=7)
    rg = ctx.align(gm, p, c, guess, neighbor_mode=7)
    assert rg["iterations"] == ro["iterations"]
    dt, dr = pose_err(ro["T"], rg["T"])
    assert dt < POSE_T_TOL and dr < POSE_R_TOL


def test_align_edge_cases(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 2)
    p, c = frames.ds[2]
    # zero correspondences: returns the guess after one "converged" iteration
    far = np.eye(4)
    far[:3, 3] = [5000.0, 5000.0, 500.0]
    rg = ctx.align(gm, p, c, far)
    ro = om.align(p, c, far)
    assert rg["iterations"] == ro["iterations"] == 1 and rg["converged"] and rg["ncorr"].tolist() == [0]
    np.testing.assert_array_equal(rg["T"], far)
    # empty cloud
    rg = ctx.align(gm, np.zeros((0, 3)), np.zeros((0, 3, 3)), far)
    assert rg["iterations"] == 1 and rg["converged"]
    np.testing.assert_array_equal(rg["T"], far)
    # max_iteration reached without convergence
    guess = frames.poses[2] @ S.perturbation()
    rg = ctx.align(gm, p, c, guess, max_iteration=2)
    ro = om.align(p, c, guess, max_iteration=2)
    assert rg["iterations"] == ro["iterations"] == 2 and not rg["converged"] and not ro["converged"]
    dt, dr = pose_err(ro["T"], rg["T"])
    assert dt < POSE_T_TOL and dr < POSE_R_TOL
    # a single point: H is rank 3, the LDLT pseudo-solve is ill-defined -> only
    # require a finite pose that stays near the guess
    rg = ctx.align(gm, p[:1], c[:1], frames.poses[2], max_iteration=5)
    assert np.isfinite(rg["T"]).all() and 1 <= rg["iterations"] <= 5


def test_device_resident_pipeline_matches_host_entry_points(ctx, oracle, frames):
    """preprocess -> align -> insert on device-resident clouds (the bench path)
    gives the same results as the host-buffer entry points."""
    om, gm = frames.build_maps(oracle, capi, ctx, 3)
    _, gm2 = frames.build_maps(oracle, capi, ctx, 3)
    xyz, t = frames.raw[3]
    raw = capi.Cloud(ctx).upload(xyz)
    ds = capi.Cloud(ctx)
    raw.preprocess_into(ds, None, frames.T_il, None, 0.5)
    gp, gc, gsrc = ds.download(want_src=True)
    hp, hc, hsrc = ctx.preprocess(xyz, t, frames.T_il, None, 0.5)
    np.testing.assert_array_equal(gp, hp)
    np.testing.assert_array_equal(gc, hc)
    np.testing.assert_array_equal(gsrc, hsrc)
    guess = frames.poses[3] @ S.perturbation()
    r1 = gm.align_cloud(ds, guess, trace=True)
    r2 = ctx.align(gm, hp, hc, guess)
    np.testing.assert_array_equal(r1["T"], r2["T"])
    assert r1["iterations"] == r2["iterations"]
    gm.insert_cloud(ds, r1["T"])
    gm2.insert(hp, hc, r2["T"])
    a, b = gm.export(), gm2.export()
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)
    # fixed-iteration variant runs exactly that many
    ds2 = capi.Cloud(ctx).upload(hp, hc)
    r3 = gm.align_cloud_fixed(ds2, guess, 7)
    assert r3["iterations"] == 7


def test_align_begin_end_and_result_paths(ctx, oracle, frames):
    """align in two halves equals align; results through host-mapped words equal the copy path;
    misuse (two registrations in flight, end without begin) is an error, not a hang."""
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    cl = capi.Cloud(ctx).upload(p, c)
    ref = gm.align_cloud(cl, guess)
    gm.align_cloud_begin(cl, guess)
    with pytest.raises(capi.EskfError):
        gm.align_cloud_begin(cl, guess)          # one registration in flight per context
    r = ctx.align_end()
    np.testing.assert_array_equal(r["T"], ref["T"])
    assert r["iterations"] == ref["iterations"] and r["converged"] == ref["converged"]
    with pytest.raises(capi.EskfError):
        ctx.align_end()
    ctx.set_option("mapped_results", 0)
    try:
        r0 = gm.align_cloud(cl, guess)
        xyz, t = frames.raw[3]
        a = ctx.preprocess(xyz, t, frames.T_il, None, 0.5)
    finally:
        ctx.set_option("mapped_results", 1)
    np.testing.assert_array_equal(r0["T"], ref["T"])
    b = ctx.preprocess(xyz, t, frames.T_il, None, 0.5)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)


def test_sharded_align_single_rank_equals_align(ctx, oracle, frames):
    om, gm = frames.build_maps(oracle, capi, ctx, 4)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    cl = capi.Cloud(ctx).upload(p, c)
    r1 = gm.align_cloud(cl, guess, trace=True)
    r2 = gm.align_cloud_sharded(cl, guess, None, trace=True)
    assert r1["iterations"] == r2["iterations"]
    dt, dr = pose_err(r1["T"], r2["T"])
    assert dt < 1e-9 and dr < 1e-9


def test_map_compact_keeps_contents(ctx, oracle, frames):
    """eskf_map_compact re-hashes into 2 x size() slots: same voxels, same statistics, and a
    registration against the compacted table gives the bit-identical pose."""
    om = oracle.Map(0.5, 1000)
    gm = capi.Map(ctx, 0.5, 1000, 1 << 18)        # far too large a table for ~1e4 voxels
    for (p, c), T in zip(frames.ds[:4], frames.poses[:4]):
        om.update(p, c, T, initialize=True)
        gm.insert(p, c, T)
    p, c = frames.ds[4]
    guess = frames.poses[4] @ S.perturbation()
    before = gm.export()
    r0 = ctx.align(gm, p, c, guess)
    cap0 = gm.capacity()
    gm.compact()
    assert gm.capacity() < cap0
    assert gm.capacity() == max(1024, gm.size()) * 2 + (-(max(1024, gm.size()) * 2)) % 64
    assert gm.size() == om.size()
    for u, v in zip(before, gm.export()):
        np.testing.assert_array_equal(u, v)
    r1 = ctx.align(gm, p, c, guess)
    np.testing.assert_array_equal(r1["T"], r0["T"])
    np.testing.assert_array_equal(r1["ncorr"], r0["ncorr"])
    gm.compact()                                   # already compact: a no-op
    assert gm.capacity() == max(1024, gm.size()) * 2 + (-(max(1024, gm.size()) * 2)) % 64
    # the compacted table keeps accepting inserts (it grows again)
    gm.insert(p, c, frames.poses[4])
    om.update(p, c, frames.poses[4], initialize=True)
    assert gm.size() == om.size()
    np.testing.assert_array_equal(gm.export()[0], om.export()[0])


def test_align_large_cloud_cta_shapes_and_ticket_chunks(oracle):
    """A cloud large enough for the load-balanced (ticketed) tail of a pass, at every CTA shape,
    load rotation depth and ticket size: per-iteration correspondence counts identical to the oracle's, H/b and the
    pose within the bars, whichever warps end up summing which tiles."""
    c2 = capi.Context(0)
    rng = np.random.default_rng(5)
    scene = S.block_scene()
    mp, mc = S.dense_cloud(scene, 2_000_000, rng)
    p, c = S.dense_cloud(scene, 1_000_000, rng)
    om = oracle.Map(0.5, 1000)
    om.update(mp, mc, np.eye(4), initialize=True)
    gm = capi.Map(c2, 0.5, 1000, 1 << 20)
    gm.insert(mp, mc, np.eye(4))
    gm.compact()
    assert gm.size() == om.size()
    guess = S.perturbation(dt=(0.03, -0.015, 0.01), angle_deg=0.3)
    ro = om.align(p, c, guess)
    cl = capi.Cloud(c2, len(p)).upload(p, c)
    try:
        # (block, depth, chunk, dynamic tiles, resident tiles per warp, flagged-word pose broadcast)
        for block, depth, chunk, dyn, res, ll in [
                (256, 3, 1, 1, -1, 1), (256, 3, 2, 1, -1, 1), (257, 0, 2, 1, -1, 1), (257, 0, 1, 0, -1, 0), (769, 0, 2, 1, -1, 1), (384, 3, 4, 1, -1, 1), (768, 3, 1, 1, -1, 1),
                (768, 3, 2, 1, -1, 1), (768, 3, 4, 1, -1, 1), (768, 3, 2, 0, -1, 1),
                (768, 4, 2, 1, -1, 1), (768, 4, 1, 0, -1, 1), (640, 4, 2, 1, -1, 1), (640, 4, 4, 0, -1, 1),
                (512, 4, 2, 1, -1, 1), (512, 4, 1, 1, -1, 1),
                # depth 5: SM-resident positions + probe filter + bulk-copy ring (the default for a cloud
                # this size): everything resident, nothing resident (all tiles through the ring, ticketed
                # tail), a few tiles resident, both CTA shapes, both pose broadcasts
                (0, 0, 2, 1, -1, 1), (640, 5, 2, 1, -1, 0), (640, 5, 2, 1, 0, 1), (640, 5, 1, 1, 0, 0),
                (640, 5, 4, 1, 3, 1), (640, 5, 2, 0, 3, 1), (640, 5, 2, 0, 0, 1), (512, 5, 2, 1, -1, 1),
                (512, 5, 2, 1, 2, 0),
                # depth 6: producer / consumer warps around a shared-memory hit queue (the default for a
                # cloud this size): everything resident / nothing / a few tiles, both CTA shapes
                (640, 6, 2, 1, -1, 1), (640, 6, 2, 1, 0, 1), (640, 6, 2, 1, 5, 0), (512, 6, 2, 1, -1, 1),
                (512, 6, 2, 1, 0, 0),
                # depth 7: phase-split passes around a per-CTA hit list: all of it in shared memory, a list
                # of 64 / 1 batch(es) in shared memory and the rest in the HBM spill region, both CTA shapes
                (640, 7, 2, 1, -1, 1), (640, 7, 2, 1, 64, 1), (640, 7, 2, 1, 1, 0), (512, 7, 2, 1, -1, 1),
                (512, 7, 2, 1, 2, 0),
                # depth 8: depth 4's register pipeline with the candidates parked per warp until 32 are there
                (640, 8, 2, 1, -1, 1), (640, 8, 1, 0, -1, 0), (512, 8, 4, 1, -1, 1), (768, 8, 2, 1, -1, 1),
                # depth 9: depth 8 with the positions streamed through a cp.async.bulk ring
                (512, 9, 2, 1, -1, 1), (512, 9, 1, 0, -1, 0), (640, 9, 4, 1, -1, 1),
                # depth 10: three launches per iteration (high-occupancy lookup -> hit list in HBM -> dense gather -> solve)
                (0, 10, 2, 1, -1, 1),
                # depth 11: parked candidates with 2 tiles of lookups per trip
                (512, 11, 2, 1, -1, 1), (384, 11, 1, 0, -1, 0)]:
            c2.set_option("align_block", block)
            c2.set_option("align_depth", depth)
            c2.set_option("align_ticket_chunk", chunk)
            c2.set_option("align_dynamic_tiles", dyn)
            c2.set_option("align_resident", res)
            c2.set_option("align_ll", ll)
            rg = gm.align_cloud(cl, guess, trace=True)
            tag = (block, depth, chunk, dyn, res, ll)
            assert rg["converged"] and rg["iterations"] == ro["iterations"], tag
            np.testing.assert_array_equal(rg["ncorr"], ro["ncorr"], err_msg=str(tag))
            for k in range(ro["iterations"]):
                assert rel_err(rg["H"][k], ro["H"][k]) < H_TOL, (tag, k)
                assert b_rel(rg["b"][k], ro["b"][k], ro["H"][k]) < H_TOL, (tag, k)
            dt, dr = pose_err(ro["T"], rg["T"])
            assert dt < POSE_T_TOL and dr < POSE_R_TOL, (tag, dt, dr)
            # the result that comes back through the host-mapped words (no trace requested)
            rq = gm.align_cloud(cl, guess)
            assert rq["iterations"] == ro["iterations"], tag
            dt, dr = pose_err(ro["T"], rq["T"])
            assert dt < POSE_T_TOL and dr < POSE_R_TOL, (tag, dt, dr)
        # the (0, 0) cell went through the context's one-off timing of the two large-cloud loop shapes
        assert c2.get_option("align_tuned_block") in (256, 257, 512, 769)
        for block, depth in ((0, 0),):
            c2.set_option("align_block", block)
            c2.set_option("align_depth", depth)
            c2.set_option("align_resident", -1)
            for tune in (0, 1, 1):
                c2.set_option("align_autotune", tune)   # (setting it forgets the earlier choice)
                assert c2.get_option("align_tuned_block") == 0
                rq = gm.align_cloud(cl, guess, trace=True)
                assert c2.get_option("align_tuned_block") == (0 if tune == 0 else c2.get_option("align_tuned_block"))
                assert (c2.get_option("align_tuned_block") in (256, 257, 512, 769)) == bool(tune)
                np.testing.assert_array_equal(rq["ncorr"], ro["ncorr"])
                dt, dr = pose_err(ro["T"], rq["T"])
                assert dt < POSE_T_TOL and dr < POSE_R_TOL, (tune, dt, dr)
        # the per-point correspondence flags of a cloud this size (depth 5, the 4-deep loop, 256-thread CTAs)
        for block, depth, res in ((0, 0, -1), (257, 0, -1), (0, 10, -1), (512, 9, -1), (640, 8, -1), (768, 8, -1), (640, 7, 3), (640, 6, 0), (640, 5, 0), (640, 5, -1), (640, 4, -1), (256, 3, -1)):
            c2.set_option("align_block", block)
            c2.set_option("align_depth", depth)
            c2.set_option("align_resident", res)
            Ho, bo, hito, nco = om.linearize(*oracle.transform_cloud(p, c, guess))
            Hg, bg, hitg, ncg = c2.linearize(gm, p, c, T=guess)
            np.testing.assert_array_equal(hitg, hito)
            assert ncg == nco == int(ro["ncorr"][0])
            assert rel_err(Hg, Ho) < H_TOL and b_rel(bg, bo, Ho) < H_TOL
        # the probe filter follows the table: insert more points, evict, compact -> still the oracle's result
        c2.set_option("align_block", 0)
        c2.set_option("align_depth", 0)
        c2.set_option("align_resident", -1)
        mp2, mc2 = S.dense_cloud(scene, 300_000, rng)
        gm.insert(mp2, mc2, np.eye(4))
        om.update(mp2, mc2, np.eye(4), initialize=True)
        for step in ("insert", "evict", "compact"):
            if step == "evict":
                assert gm.evict(np.zeros(3), 60.0) == om.evict(np.zeros(3), 60.0)
            if step == "compact":
                gm.compact()
            ro2 = om.align(p, c, guess)
            rg2 = gm.align_cloud(cl, guess, trace=True)
            assert rg2["iterations"] == ro2["iterations"], step
            np.testing.assert_array_equal(rg2["ncorr"], ro2["ncorr"], err_msg=step)
            dt, dr = pose_err(ro2["T"], rg2["T"])
            assert dt < POSE_T_TOL and dr < POSE_R_TOL, (step, dt, dr)
        with pytest.raises(capi.EskfError):
            c2.set_option("align_block", 500)
        with pytest.raises(capi.EskfError):
            c2.set_option("align_depth", 12)
        with pytest.raises(capi.EskfError):
            c2.set_option("align_ticket_chunk", 3)
    finally:
        c2.set_option("align_block", 0)
        c2.set_option("align_depth", 0)
        c2.set_option("align_ticket_chunk", 2)
        c2.set_option("align_dynamic_tiles", 1)
        c2.set_option("align_resident", -1)
        c2.set_option("align_ll", 1)


def test_cuda_path_against_the_reference_sources(ctx, oracle, frames):
    """The CUDA path held DIRECTLY against the reference's own sources (compiled on the API shims,
    oracle/ref.py): kept sets with deskew, map statistics, correspondence set bit for bit, final
    pose within the bar.  The prebuilt oracle/_ref libraries travel with the snapshot."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref libraries not present on this box")
    from ref_check import GpuImpl, check_against_reference
    from test_reference_shim import sweep_and_states
    check_against_reference(R.Ref("seq"), GpuImpl(capi, ctx), frames, sweep_and_states(7, 2),
                            cov_tol=1e-6, pose_tol=POSE_T_TOL)


def test_align_batch_equals_individual_registrations(oracle, frames):
    """eskf_align_batch (configs[4]): 7 jobs over 3 contexts, one in flight per context, give the
    bit-identical poses and iteration counts of 7 separate eskf_align_cloud calls (and the oracle's
    pose within the bar); a job on the wrong context is an error that leaves the contexts usable."""
    ctxs = [capi.Context(0) for _ in range(3)]
    jobs, want = [], []
    for i in range(7):
        c = ctxs[i % 3]
        om = oracle.Map(0.5, 1000)
        gm = capi.Map(c, 0.5, 1000, 1 << 14)
        for (p, cv), T in zip(frames.ds[:3 + i % 2], frames.poses[:3 + i % 2]):
            om.update(p, cv, T, initialize=True)
            gm.insert(p, cv, T)
        p, cv = frames.ds[4]
        f = 0.4 + 0.1 * i                                   # scaled config-1 perturbations: all converge
        guess = frames.poses[4] @ S.perturbation(dt=(0.10 * f, -0.05 * f, 0.03 * f), angle_deg=f)
        cl = capi.Cloud(c, len(p)).upload(p, cv)
        jobs.append((gm, cl, guess))
        want.append((gm.align_cloud(cl, guess), om.align(p, cv, guess)))
    res = capi.align_batch(ctxs, [j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs])
    assert len(res) == 7
    for r, (single, ro) in zip(res, want):
        np.testing.assert_array_equal(r["T"], single["T"])
        assert r["iterations"] == single["iterations"] == ro["iterations"]
        assert r["converged"] == single["converged"] == bool(ro["converged"])
        dt, dr = pose_err(ro["T"], r["T"])
        assert dt < POSE_T_TOL and dr < POSE_R_TOL
    assert capi.align_batch(ctxs, [], [], []) == []
    # job 1 handed a cloud of context 0: refused; job 0 (in flight then) is still collected
    bad_clouds = [jobs[0][1], jobs[0][1], jobs[2][1]]
    with pytest.raises(capi.EskfError):
        capi.align_batch(ctxs, [j[0] for j in jobs[:3]], bad_clouds, [j[2] for j in jobs[:3]])
    again = capi.align_batch(ctxs, [j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs])
    for r, (single, _) in zip(again, want):
        np.testing.assert_array_equal(r["T"], single["T"])
