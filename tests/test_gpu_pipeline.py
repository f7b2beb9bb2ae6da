"""End-to-end parity of the C++ orchestrator + CUDA library against the oracle pipeline (Removerter::run())."""
import numpy as np
import pytest

import oracle
from lt_mapper_b200 import removert

pytestmark = pytest.mark.gpu

MAPS = ["map_global_orig_", "map_global_curr_", "map_global_curr_static_", "map_global_curr_dynamic_", "map_global_nd_",
        "map_global_nd_strong_", "map_global_nd_weak_", "map_global_pd_", "map_global_pd_orig_", "map_global_pd_strong_",
        "map_global_pd_weak_"]
SAVED = ["OriginalNoisyCentralMapGlobal", "OriginalNoisyQueryMapGlobal", "central_sess_high_dyn", "query_sess_high_dyn",
         "union_map_queryside", "union_map_centralside", "pd_map", "nd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map",
         "weak_pd_map"]
SCANSETS = ["keyframe_scans_", "keyframe_scans_static_projected_", "keyframe_scans_dynamic_", "scans_knn_coexist_", "scans_knn_diff_"]


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _run_both(pair, schedule, num_knn, knn_thr, step3=False, **kw):
    c, q = pair
    O = oracle.Removerter(num_knn=num_knn, knn_thr=knn_thr, schedule=schedule)
    G = removert.Removerter(num_knn=num_knn, knn_thr=knn_thr, schedule=schedule, **kw)
    for s, d in ((0, c), (1, q)):
        inv = oracle.inverse_poses(d.poses)
        O.load_session(s, d.xyzi, d.offsets, d.poses, inv)
        G.load_session(s, d.xyzi, d.offsets, d.poses, inv)
    O.run(step0=True, step12=True, step3=step3)
    G.run_step0(); G.run_step12()
    if step3:
        G.run_step3()
    return O, G


def _compare(O, G, step3=False):
    assert O.log() == G.log()
    for s in (0, 1):
        for n in MAPS:
            try:
                g = G.cloud(n, s)
            except Exception:
                g = np.zeros((0, 4), np.float32)
            assert _same(O.cloud(n, s), g), (n, s)
        for n in SCANSETS:
            pts, off = G.scanset(n, s)
            for k, e in enumerate(O.clouds(n, s)):
                assert _same(e, pts[off[k]:off[k + 1]]), (n, s, k)
    for n in SAVED + (["updated_map", "updated_map_strong"] if step3 else []):
        assert _same(O.cloud("saved:" + n), G.cloud("saved:" + n)), n
    if step3:
        for n in ["keyframe_scans_updated_", "keyframe_scans_updated_strong_", "keyframe_scans_pd_", "keyframe_scans_strong_pd_",
                  "keyframe_scans_strong_nd_", "keyframe_scans_weak_nd_"]:
            pts, off = G.scanset(n, 0)
            for k, e in enumerate(O.clouds(n, 0)):
                assert _same(e, pts[off[k]:off[k + 1]]), (n, k)


def test_shipped_run(small_pair):
    """Shipped run(): one removeOnce(2.5) per session, yaml kNN (k=2, 0.01), Steps 0-3."""
    O, G = _run_both(small_pair, [(0, 2.5)], 2, 0.01, step3=True)
    _compare(O, G, step3=True)
    assert O.log()[0][2] > 0
    G.close()


def test_multires_schedule(small_pair):
    """BASELINE config 2 shape: 3 remove + 1 revert resolutions, kNN k=1 r=0.2 m (threshold r^2 = 0.04)."""
    sched = [(0, 2.5), (0, 2.0), (0, 1.5), (1, 1.0)]
    O, G = _run_both(small_pair, sched, 1, 0.04, keyframe_batch=4)
    _compare(O, G)
    G.close()


def test_selfremovert_schedule(small_pair):
    sched = removert.selfremovert_schedule([2.5, 2.0])
    O, G = _run_both(small_pair, sched, 2, 0.01)
    _compare(O, G)
    G.close()


def test_in_memory_cascade(small_pair):
    """Three-session LT-map cascade through ltrh_cascade_promote_updated == the oracle chained by hand the way the reference's file
    protocol chains runs (scans_updated -> VoxelGrid at load -> next central session).  The file-level equivalent is checked
    against the compiled reference in tests/test_gpu_driver.py."""
    import synth
    K = 4
    sess = [synth.make_session(s, K, beams=32, az_steps=900) for s in (0, 1, 2)]

    def grid(d):   # Session::loadKeyframes' pcl::VoxelGrid (Session.cpp:284-289) on every scan
        sc = [removert.voxel_grid(d.scan(k), 0.05)[0] for k in range(d.K)]
        return np.concatenate(sc), np.concatenate([[0], np.cumsum([len(a) for a in sc])]).astype(np.int64)

    loaded = [grid(d) for d in sess]
    inv = [oracle.inverse_poses(d.poses) for d in sess]
    # oracle, chained by hand
    O1 = oracle.Removerter(num_knn=2, knn_thr=0.01)
    O1.load_session(0, *loaded[0], sess[0].poses, inv[0]); O1.load_session(1, *loaded[1], sess[1].poses, inv[1])
    O1.run(step3=True)
    upd = [removert.voxel_grid(a, 0.05)[0] for a in O1.clouds("keyframe_scans_updated_", 0)]
    O2 = oracle.Removerter(num_knn=2, knn_thr=0.01)
    O2.load_session(0, np.concatenate(upd), np.concatenate([[0], np.cumsum([len(a) for a in upd])]).astype(np.int64), sess[0].poses, inv[0])
    O2.load_session(1, *loaded[2], sess[2].poses, inv[2])
    O2.run(step3=True)
    # device, promoted in place
    G = removert.Removerter(num_knn=2, knn_thr=0.01)
    G.load_session(0, *loaded[0], sess[0].poses, inv[0]); G.load_session(1, *loaded[1], sess[1].poses, inv[1])
    G.run_step0(); G.run_step12(); G.run_step3()
    _compare(O1, G, step3=True)
    G.cascade_promote_updated()
    pts, off = G.scanset("keyframe_scans_", 0)
    assert all(_same(pts[off[k]:off[k + 1]], upd[k]) for k in range(K))
    G.load_session(1, *loaded[2], sess[2].poses, inv[2])
    G.run_step0(); G.run_step12(); G.run_step3()
    _compare(O2, G, step3=True)
    assert len(O2.cloud("saved:nd_map")) > 0 and len(O2.cloud("saved:pd_map")) > 0
    with pytest.raises(Exception):
        G2 = removert.Removerter()
        try:
            G2.cascade_promote_updated()      # Step 3 has not run
        finally:
            G2.close()
    G.close()
