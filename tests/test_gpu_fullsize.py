"""BASELINE configs[1] size (200-keyframe pair, 64x1800 scans, ~7 M-point maps): properties that do not need the oracle at
full size, plus oracle spot checks on a few keyframes against the full map."""
import numpy as np
import pytest

import oracle
import lt_mapper_b200 as ltr
from lt_mapper_b200 import removert

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import synth
    c, q = synth.make_pair(200)
    R = removert.Removerter(num_knn=1, knn_thr=0.04, schedule=[(0, 2.5), (0, 2.0), (0, 1.5), (1, 1.0)])
    for s, d in ((0, c), (1, q)):
        R.load_session(s, d.xyzi, d.offsets, d.poses, np.stack([np.linalg.inv(p) for p in d.poses]))
    R.run_step0()
    yield c, q, R
    R.close()


def test_fast_equals_exact_and_oracle_spot_checks(big):
    c, q, R = big
    m = R.cloud("map_global_curr_", 0)
    assert len(m) > 5_000_000
    inv = np.stack([np.linalg.inv(p) for p in c.poses])
    flags = {}
    for fast in (True, "nocull", False):
        with ltr.Context(fast_path=bool(fast), cull=(fast is True)) as ctx:
            mh = ctx.cloud_upload(m); ss = ctx.scanset_upload(c.xyzi, c.offsets); ps = ctx.poses_upload(c.poses, inv)
            n = ctx.remove_pass(mh, ss, ps, ltr.MODE_HD, 2.5)
            f = ctx.flags_download(mh)
            assert n == int(f.sum())
            # oracle spot check: keyframes 17..19 alone against the full map
            ctx.remove_pass(mh, ss, ps, ltr.MODE_HD, 2.5, kf_begin=17, kf_end=20)
            sub = ctx.flags_download(mh)
            st, dy = ctx.apply_partition(mh)
            assert ctx.cloud_size(st) + ctx.cloud_size(dy) == len(m) and ctx.cloud_size(dy) == int(sub.sum())
            vis = ctx.parse_projected(mh, ps, 100, 102, 3.0)
            flags[fast] = (f, sub, ctx.scanset_download(vis))
    assert np.array_equal(flags[True][0], flags[False][0]) and np.array_equal(flags["nocull"][0], flags[False][0])
    assert np.array_equal(flags[True][1], flags[False][1]) and np.array_equal(flags["nocull"][1], flags[False][1])
    assert np.array_equal(flags[True][2][0].view(np.uint32), flags[False][2][0].view(np.uint32)) and np.array_equal(flags[True][2][1], flags[False][2][1])
    sl = slice(c.offsets[17], c.offsets[20])
    exp = oracle.remove_pass(m, c.xyzi[sl], c.offsets[17:21] - c.offsets[17], inv[17:20], oracle.MODE_HD, 2.5)
    assert np.array_equal(flags[True][1], exp)
    # union over keyframes is monotone: the 3-keyframe flags are a subset of the 200-keyframe flags
    assert not np.any(flags[True][1] & ~flags[True][0])
    pts, off = flags[True][2]
    e, _ = oracle.parse_projected(m, inv[100], 3.0)
    assert np.array_equal(pts[off[0]:off[1]].view(np.uint32), e.view(np.uint32))
    assert (np.diff(off) <= 150 * 1080).all()


def test_pipeline_invariants_at_full_size(big):
    c, q, R = big
    R.reset_to_step0(); R.run_step12()
    log = R.log()
    assert [l[0] for l in log[:8]] == ["removeOnce", "removeOnce", "removeOnce", "revertOnce"] * 2
    for what, n_map, n_dyn, n_static_after, n_dyn_after in log:
        assert 0 <= n_dyn <= n_map and n_static_after >= 0
        if what == "removeOnce":
            assert n_static_after <= n_map - n_dyn           # re-voxelising the survivors never adds points
    # every ND / PD change point came out of the kNN diff of a visible point: counts are consistent
    for s in (0, 1):
        vis = R.scanset("keyframe_scans_static_projected_", s)[1]
        co = R.scanset("scans_knn_coexist_", s)[1]; di = R.scanset("scans_knn_diff_", s)[1]
        assert np.array_equal(np.diff(vis), np.diff(co) + np.diff(di))
    nd, snd, wnd = (R.cloud_size("saved:" + n) for n in ("nd_map", "strong_nd_map", "weak_nd_map"))
    assert 0 < snd and 0 < wnd and snd + wnd <= nd * 1.01 + 10
    # determinism: a second run gives byte-identical change maps
    a = {n: R.cloud("saved:" + n) for n in ("strong_nd_map", "weak_pd_map", "pd_map")}
    R.reset_to_step0(); R.run_step12()
    for n, v in a.items():
        assert np.array_equal(v.view(np.uint32), R.cloud("saved:" + n).view(np.uint32))
    assert R.log() == log
