"""Committed golden vectors (tests/golden/small_pair.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (regression guard).  GPU: the C-ABI path reproduces them without the oracle."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "small_pair.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_oracle_reproduces_golden(gold):
    import oracle
    g = gold
    merged = np.concatenate([oracle.transform(g["c_xyzi"][g["c_off"][k]:g["c_off"][k + 1]], g["c_poses"][k]) for k in range(3)])
    m = oracle.voxel(merged, 0.05)
    assert np.array_equal(_bits(m), _bits(g["map"]))
    assert np.array_equal(_bits(oracle.voxel(merged, 0.4)), _bits(g["vox_04"]))
    for mode, nm in ((oracle.MODE_HD, "hd"), (oracle.MODE_ND, "nd"), (oracle.MODE_PD, "pd")):
        for a in (2.5, 1.5):
            f = oracle.remove_pass(m, g["q_xyzi"], g["q_off"], g["q_inv"], mode, a)
            assert np.array_equal(np.packbits(f), g[f"flags_{nm}_{a}"])
    pts, idx = oracle.parse_projected(m, g["c_inv"][1], 3.0)
    assert np.array_equal(idx, g["vis_idx"]) and np.array_equal(_bits(pts), _bits(g["vis_pts"]))
    lab, co, di = oracle.knn_partition(g["c_xyzi"][:g["c_off"][1]], g["c_poses"][0], g["c_inv"][0], m[::3], 2, 0.01)
    assert np.array_equal(np.packbits(lab), g["knn_lab"]) and np.array_equal(_bits(co), _bits(g["knn_co"])) and np.array_equal(_bits(di), _bits(g["knn_di"]))


@pytest.mark.gpu
def test_gpu_reproduces_golden(gold):
    import lt_mapper_b200 as ltr
    from lt_mapper_b200 import removert
    g = gold
    with ltr.Context() as ctx:
        cs = ctx.scanset_upload(g["c_xyzi"], g["c_off"]); cp = ctx.poses_upload(g["c_poses"], g["c_inv"])
        qs = ctx.scanset_upload(g["q_xyzi"], g["q_off"]); qp = ctx.poses_upload(g["q_poses"], g["q_inv"])
        merged = ctx.merge_scans_global(cs, cp)
        mh = ctx.voxel_centroid(merged, 0.05)
        assert np.array_equal(_bits(ctx.cloud_download(mh)), _bits(g["map"]))
        assert np.array_equal(_bits(ctx.cloud_download(ctx.voxel_centroid(merged, 0.4))), _bits(g["vox_04"]))
        for mode, nm in ((ltr.MODE_HD, "hd"), (ltr.MODE_ND, "nd"), (ltr.MODE_PD, "pd")):
            for a in (2.5, 1.5):
                ctx.remove_pass(mh, qs, qp, mode, a)
                assert np.array_equal(np.packbits(ctx.flags_download(mh)), g[f"flags_{nm}_{a}"]), (nm, a)
        vis = ctx.parse_projected(mh, cp, 1, 2, 3.0)
        pts, off = ctx.scanset_download(vis)
        assert np.array_equal(_bits(pts), _bits(g["vis_pts"]))
        th = ctx.cloud_upload(g["map"][::3])
        s0 = ctx.scanset_upload(g["c_xyzi"][:g["c_off"][1]], g["c_off"][:2])
        co, di = ctx.knn_diff(s0, cp, th, 2, 0.01)
        assert np.array_equal(_bits(ctx.scanset_download(co)[0]), _bits(g["knn_co"]))
        assert np.array_equal(_bits(ctx.scanset_download(di)[0]), _bits(g["knn_di"]))
    R = removert.Removerter(num_knn=2, knn_thr=0.01, schedule=[(0, 2.5), (1, 2.375), (0, 2.5)])
    R.load_session(0, g["c_xyzi"], g["c_off"], g["c_poses"], g["c_inv"])
    R.load_session(1, g["q_xyzi"], g["q_off"], g["q_poses"], g["q_inv"])
    R.run_step0(); R.run_step12(); R.run_step3()
    assert np.array_equal(np.array([l[1:] for l in R.log()], np.int64), g["pipe_log"])
    for n in ("nd_map", "pd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "updated_map"):
        assert np.array_equal(_bits(R.cloud("saved:" + n)), _bits(g["pipe_" + n])), n
    R.close()
