"""GPU parity tests: every C-ABI hot-path entry point against the CPU oracle on identical inputs.
Bar: bit-exact indices / flags / labels / coordinates (the arithmetic is restated operation by operation)."""
import numpy as np
import pytest

import oracle
import lt_mapper_b200 as ltr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ltr.Context()
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_device_math_bit_exact(ctx):
    """cart2sph + pixel index (utility.cpp:38-56, 118-123) on 4M random points + edge cases, all resolutions."""
    rng = np.random.default_rng(7)
    n = 4_000_000
    xyz = np.empty((n, 3), np.float32)
    xyz[:, 0] = rng.uniform(-100, 100, n); xyz[:, 1] = rng.uniform(-100, 100, n); xyz[:, 2] = rng.uniform(-10, 10, n)
    edge = np.array([[10, 0, 0], [0, 10, 0], [-10, 0, 0], [-10, -1e-6, 0], [0, 0, 0], [1, 1, 1], [-1, -1e-30, 0], [0, 0, 5],
                     [0, 0, -5], [1e-20, 1e-20, 1e-20], [1, 0, 0], [-1, 0, 0], [3, 4, 0], [1, 0, 0.4375], [1, 0, 2.4375]], np.float32)
    xyz[:len(edge)] = edge
    # near-sensor and grazing points
    xyz[100:100000] *= rng.uniform(0.001, 0.1, (99900, 1)).astype(np.float32)
    # log-uniform magnitudes 2^-30 .. 2^30 per coordinate: ratios around atanf's branch thresholds (|y/x| >= 2^25, < 2^-29)
    m = 200000
    wide = (np.exp2(rng.uniform(-30, 30, (m, 3))) * rng.choice([-1.0, 1.0], (m, 3))).astype(np.float32)
    wide[:4] = [[0.0078125, 340000.0, 0.0], [1.0, 2.0**25, 0.0], [1.0, 2.0**25 - 2, 0.0], [0.0, 0.0078125, 340000.0]]
    xyz[100000:100000 + m] = wide
    for alpha in (2.5, 3.0, 1.425):
        rows, cols = oracle.reset_rimg_size(alpha)
        assert (rows, cols) == ltr.reset_rimg_size(alpha)
        r0, c0, rg0 = oracle.pixel_index(xyz, rows, cols)
        r1, c1, rg1, az1, el1 = ctx.debug_pixel_index(xyz, rows, cols)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1)
        assert np.array_equal(_bits(rg0), _bits(rg1))
    az0 = oracle.atan2f(xyz[:, 1], xyz[:, 0])
    assert np.array_equal(_bits(az0), _bits(az1))
    el0 = oracle.atan2f(xyz[:, 2], np.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]).astype(np.float32), dtype=np.float32))
    assert np.array_equal(_bits(el0), _bits(el1))


@pytest.mark.parametrize("order", [0, 1])
def test_merge_and_voxel(small_pair, order):
    c = small_pair[0]
    with ltr.Context(transform_order=order) as ctx:
        ss = ctx.scanset_upload(c.xyzi, c.offsets)
        inv = oracle.inverse_poses(c.poses)
        ps = ctx.poses_upload(c.poses, inv)
        merged = ctx.merge_scans_global(ss, ps)
        got = ctx.cloud_download(merged)
        exp = np.concatenate([oracle.transform(c.scan(k), c.poses[k], order) for k in range(c.K)])
        assert np.array_equal(_bits(got), _bits(exp))
        for leaf in (0.05, 0.4):
            v = ctx.voxel_centroid(merged, leaf)
            gv = ctx.cloud_download(v)
            ev = oracle.voxel(exp, leaf)
            assert gv.shape == ev.shape
            assert np.array_equal(_bits(gv), _bits(ev))


@pytest.mark.parametrize("mode,alpha", [(ltr.MODE_HD, 2.5), (ltr.MODE_HD, 2.0), (ltr.MODE_HD, 1.425), (ltr.MODE_PD, 2.5), (ltr.MODE_ND, 2.5)])
@pytest.mark.parametrize("batch", [4, 32])
def test_remove_pass_flags(small_pair, small_maps, mode, alpha, batch):
    c = small_pair[0]
    m = small_maps[0] if mode == ltr.MODE_HD else small_maps[1][::7]  # cross-session variants project the other map
    inv = oracle.inverse_poses(c.poses)
    exp = oracle.remove_pass(m, c.xyzi, c.offsets, inv, mode, alpha)
    with ltr.Context(keyframe_batch=batch) as ctx:
        mh = ctx.cloud_upload(m)
        ss = ctx.scanset_upload(c.xyzi, c.offsets)
        ps = ctx.poses_upload(c.poses, inv)
        n = ctx.remove_pass(mh, ss, ps, mode, alpha)
        got = ctx.flags_download(mh)
        assert n == int(exp.sum())
        assert np.array_equal(got, exp)
        # keyframe-sharded accumulation == single call
        ctx.remove_pass(mh, ss, ps, mode, alpha, kf_begin=0, kf_end=c.K // 2)
        n2 = ctx.remove_pass(mh, ss, ps, mode, alpha, kf_begin=c.K // 2, kf_end=c.K, accumulate=True)
        assert n2 == n and np.array_equal(ctx.flags_download(mh), exp)
        st, dy = ctx.apply_partition(mh)
        assert np.array_equal(_bits(ctx.cloud_download(st)), _bits(m[exp == 0]))
        assert np.array_equal(_bits(ctx.cloud_download(dy)), _bits(m[exp == 1]))
    assert exp.sum() > 0


def test_parse_projected(small_pair, small_maps, ctx):
    c = small_pair[0]
    m = small_maps[0]
    inv = oracle.inverse_poses(c.poses)
    mh = ctx.cloud_upload(m)
    ps = ctx.poses_upload(c.poses, inv)
    out = ctx.parse_projected(mh, ps, 1, c.K, 3.0)
    pts, off = ctx.scanset_download(out)
    assert len(off) == c.K
    for k in range(1, c.K):
        e, _ = oracle.parse_projected(m, inv[k], 3.0)
        g = pts[off[k - 1]:off[k]]
        assert g.shape == e.shape, k
        assert np.array_equal(_bits(g), _bits(e)), k


@pytest.mark.parametrize("k,thr", [(2, 0.01), (1, 0.04), (3, 0.1)])
def test_knn_diff(small_pair, small_maps, ctx, k, thr):
    c = small_pair[0]
    target = small_maps[1]
    inv = oracle.inverse_poses(c.poses)
    th = ctx.cloud_upload(target)
    ps = ctx.poses_upload(c.poses, inv)
    ss = ctx.scanset_upload(c.xyzi, c.offsets)
    co, di = ctx.knn_diff(ss, ps, th, k, thr)
    cop, coo = ctx.scanset_download(co)
    dip, dio = ctx.scanset_download(di)
    for kf in range(c.K):
        lab, eco, edi = oracle.knn_partition(c.scan(kf), c.poses[kf], inv[kf], target, k, thr)
        assert np.array_equal(_bits(cop[coo[kf]:coo[kf + 1]]), _bits(eco)), kf
        assert np.array_equal(_bits(dip[dio[kf]:dio[kf + 1]]), _bits(edi)), kf
        assert 0 < lab.sum() < len(lab)


def test_voxel_shortcut_for_one_point_per_voxel_inputs():
    """A voxelised cloud (>= 100k points), and a subset that keeps its bounding box, re-voxelised at the same leaf: the sort is skipped
    (ltr_voxel_shortcuts counts it) and the result still equals the oracle's; a subset that moves the box takes the normal path."""
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-40, 40, (600000, 3)), rng.uniform(0, 1, (600000, 1))], 1).astype(np.float32)
    pts[0, :3] = -0.0                                                    # 0.0f + (-0.0f) = +0.0f must survive the shortcut
    v = oracle.voxel(pts, 0.2)
    assert len(v) > 150000
    lo, hi = v[:, :3].argmin(0), v[:, :3].argmax(0)
    keep = rng.random(len(v)) < 0.8
    keep[lo] = True; keep[hi] = True                                      # extremes stay -> same box
    sub_same = v[keep]
    drop = np.ones(len(v), bool); drop[lo[0]] = False                     # the x-minimum goes -> the box (and the grid) moves
    sub_moved = v[drop]
    with ltr.Context() as ctx:
        for cloud, expect_shortcut in ((pts, 0), (v, 1), (sub_same, 1), (sub_moved, None)):
            before = ctx.voxel_shortcuts()
            got = ctx.cloud_download(ctx.voxel_centroid(ctx.cloud_upload(cloud), 0.2))
            exp = oracle.voxel(cloud, 0.2)
            assert got.shape == exp.shape and np.array_equal(got.view(np.uint32), exp.view(np.uint32))
            if expect_shortcut is not None:
                assert ctx.voxel_shortcuts() - before == expect_shortcut
