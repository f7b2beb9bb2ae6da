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
