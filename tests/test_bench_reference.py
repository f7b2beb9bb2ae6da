"""bench.py --impl reference: the sampling machinery (the reference's own per-keyframe loops, timed on maps of prescribed sizes) on a tiny pair.
CPU only; needs oracle/_ref/libltremovert_ref_omp.so (built where /root/reference is mounted; it travels with the repository snapshot)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_sampler_on_a_tiny_pair():
    import bench
    import oracle
    import synth
    from oracle import ref
    if not ref.available(omp=True):
        pytest.skip("oracle/_ref/libltremovert_ref_omp.so not built")
    K, kw = 4, dict(beams=16, az_steps=600)
    c, q = synth.make_pair(K, **kw)
    # the pass-size table of THIS pair, from the oracle's run of the bench schedule
    R = oracle.Removerter(num_knn=bench.NUM_KNN, knn_thr=bench.KNN_THR, schedule=bench.schedule(), threads=2)
    for s, d in ((0, c), (1, q)):
        R.load_session(s, d.xyzi, d.offsets, d.poses)
    R.run(step0=True, step12=True)
    log = R.log()
    hd = [e for e in log if e[0] in ("removeOnce", "revertOnce")]
    n = len(bench.schedule())
    assert len(hd) == 2 * n
    sizes = {"keyframes_per_session": K, "map_points": [hd[0][1], hd[n][1]],
             "hd_pass_map_points": [[e[1] for e in hd[:n]], [e[1] for e in hd[n:]]],
             "static_map_points": [hd[n - 1][3], hd[2 * n - 1][3]],
             "nd_pass_map_points": [e[1] for e in log if e[0] == "iremoveOnceForND"],
             "pd_pass_map_points": [e[1] for e in log if e[0] == "removeOnceForPD"]}
    # the sampler re-derives the full maps with the REFERENCE's own Step 0 and refuses to run if their sizes differ from the table
    smp = bench.ReferenceSampler(K, S=1, sizes=sizes, synth_kwargs=kw)
    assert [len(m) for m in smp.full_map] == sizes["map_points"]
    t1, n1 = smp.step()
    t2, n2 = smp.step()
    assert n1 == n2 == 2 and t1 > 0 and t2 > 0
    assert set(smp.last_breakdown) == {"hd_passes", "hd_knn", "merges", "parse_static", "ld_knn", "nd_pd_passes"}
    assert "TRUE per-pass size" in smp.describe(t2)
    # a wrong table is rejected
    bad = dict(sizes, map_points=[sizes["map_points"][0] + 1, sizes["map_points"][1]])
    with pytest.raises(RuntimeError):
        bench.ReferenceSampler(K, S=1, sizes=bad, synth_kwargs=kw)


def test_config_is_identical_in_both_arms_and_matches_the_size_table():
    import json
    import bench
    cfg = bench.workload_config(bench.KF_PER_SESSION)
    assert cfg["keyframes_per_session"] == 1000 and len(cfg["schedule"]) == 9 and cfg["num_knn"] == 2
    sizes = json.load(open(bench.SIZES_PATH))
    assert sizes["keyframes_per_session"] == bench.KF_PER_SESSION
    assert len(sizes["hd_pass_map_points"][0]) == len(bench.schedule()) == len(sizes["hd_pass_map_points"][1])
    assert bench.rimg_shape(2.5) == (125, 900) and bench.rimg_shape(1.425) == (71, 513) and bench.rimg_shape(3.0) == (150, 1080)
