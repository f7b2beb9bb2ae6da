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


# ---- golden vectors produced by the reference's own code (tests/golden/ref_small_pair.json, made by make_ref_golden.py) ----
REF_GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_small_pair.json")
REF_MAPS = ["map_global_orig_", "map_global_curr_", "map_global_curr_static_", "map_global_curr_dynamic_", "map_global_nd_", "map_global_nd_strong_",
            "map_global_nd_weak_", "map_global_pd_", "map_global_pd_orig_", "map_global_pd_strong_", "map_global_pd_weak_"]
REF_SCANSETS = ["keyframe_scans_", "keyframe_scans_static_projected_", "keyframe_scans_dynamic_", "scans_knn_coexist_", "scans_knn_diff_"]
REF_STEP3_SCANSETS = ["keyframe_scans_updated_", "keyframe_scans_updated_strong_", "keyframe_scans_pd_", "keyframe_scans_strong_pd_",
                      "keyframe_scans_strong_nd_", "keyframe_scans_weak_nd_"]
REF_SAVED = ["OriginalNoisyCentralMapGlobal", "OriginalNoisyQueryMapGlobal", "central_sess_high_dyn", "query_sess_high_dyn", "union_map_queryside",
             "union_map_centralside", "pd_map", "nd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "updated_map", "updated_map_strong"]


def _digest(a):
    import hashlib
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 4)
    return [int(a.shape[0]), hashlib.sha256(a.tobytes()).hexdigest()]


@pytest.fixture(scope="module")
def ref_gold():
    import json
    with open(REF_GOLD) as f:
        return json.load(f)


def _check_against_reference_golden(rg, cloud, scans):
    """cloud(name, sess) -> (n,4) array (empty if the implementation holds none); scans(name, sess) -> list of (n,4) arrays."""
    for s in (0, 1):
        for n in REF_MAPS + (["map_global_updated_", "map_global_updated_strong_"] if s == 0 else []):
            assert _digest(cloud(n, s)) == rg[f"map{s}:{n}"], (n, s)
        for n in REF_SCANSETS + (REF_STEP3_SCANSETS if s == 0 else []):
            assert [_digest(a) for a in scans(n, s)] == rg[f"scans{s}:{n}"], (n, s)
    for n in REF_SAVED:
        assert _digest(cloud("saved:" + n, 0)) == rg[f"saved:{n}.pcd"], n
    # the per-keyframe files saveAllTypeOfScans writes (Removerter.cpp:1606-1650) are the Step-3 scan sets
    for sub, member in (("scans_updated", "keyframe_scans_updated_"), ("scans_updated_strong", "keyframe_scans_updated_strong_"),
                        ("scans_pd", "keyframe_scans_pd_"), ("scans_pd_strong", "keyframe_scans_strong_pd_"), ("scans_nd_strong", "keyframe_scans_strong_nd_")):
        for k, d in enumerate(rg[f"scans0:{member}"]):
            assert rg[f"saved:{sub}/{k:06d}.pcd"] == d


def test_oracle_reproduces_reference_golden(gold, ref_gold):
    import oracle
    g, rg = gold, ref_gold
    P = rg["params"]
    O = oracle.Removerter(num_knn=P["num_knn"], knn_thr=P["knn_thr"], voxel=P["voxel"], order=P["order"], schedule=[(0, 2.5)])
    for s, nm in ((0, "c"), (1, "q")):
        inv = np.array(rg[f"inv{s}"], np.float64).reshape(-1, 4, 4)
        assert np.array_equal(inv, oracle.inverse_poses(g[nm + "_poses"]))          # Eigen inverse stand-in == oracle restatement
        O.load_session(s, g[nm + "_xyzi"], g[nm + "_off"], g[nm + "_poses"], inv)
    O.run(step0=True, step12=True, step3=True)

    def scans(n, s):
        try:
            return O.clouds(n, s)
        except KeyError:
            return []
    _check_against_reference_golden(rg, lambda n, s: O.cloud(n, s), scans)


@pytest.mark.gpu
def test_gpu_reproduces_reference_golden(gold, ref_gold):
    from lt_mapper_b200 import removert
    g, rg = gold, ref_gold
    P = rg["params"]
    G = removert.Removerter(num_knn=P["num_knn"], knn_thr=P["knn_thr"], voxel=P["voxel"], schedule=[(0, 2.5)])
    for s, nm in ((0, "c"), (1, "q")):
        G.load_session(s, g[nm + "_xyzi"], g[nm + "_off"], g[nm + "_poses"], np.array(rg[f"inv{s}"], np.float64).reshape(-1, 4, 4))
    G.run_step0(); G.run_step12(); G.run_step3()

    def cloud(n, s):
        try:
            return G.cloud(n, s)
        except Exception:
            return np.zeros((0, 4), np.float32)

    def scans(n, s):
        pts, off = G.scanset(n, s)
        return [pts[off[k]:off[k + 1]] for k in range(len(off) - 1)]
    _check_against_reference_golden(rg, cloud, scans)
    G.close()
