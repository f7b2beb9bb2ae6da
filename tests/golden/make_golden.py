"""Generates tests/golden/small_pair.npz: a tiny synthetic two-session pair (inputs included, so the fixture does not depend
on the generator's libm) and the oracle's outputs for every hot-path entry point.  The reference is C++ with ROS/PCL
dependencies and cannot run here, so these vectors pin the ORACLE (regression guard) and give the GPU tests a
reference-free expected value on the GPU box.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402


def main():
    c, q = synth.make_pair(3, beams=12, az_steps=400)
    out = {}
    for name, d in (("c", c), ("q", q)):
        out[name + "_xyzi"], out[name + "_off"], out[name + "_poses"] = d.xyzi, d.offsets, d.poses
        out[name + "_inv"] = oracle.inverse_poses(d.poses)
    merged = np.concatenate([oracle.transform(c.scan(k), c.poses[k]) for k in range(c.K)])
    m = oracle.voxel(merged, 0.05)
    out["map"] = m
    out["vox_04"] = oracle.voxel(merged, 0.4)
    for mode, nm in ((oracle.MODE_HD, "hd"), (oracle.MODE_ND, "nd"), (oracle.MODE_PD, "pd")):
        for a in (2.5, 1.5):
            out[f"flags_{nm}_{a}"] = np.packbits(oracle.remove_pass(m, q.xyzi, q.offsets, out["q_inv"], mode, a))
    pts, idx = oracle.parse_projected(m, out["c_inv"][1], 3.0)
    out["vis_pts"], out["vis_idx"] = pts, idx
    lab, co, di = oracle.knn_partition(c.scan(0), c.poses[0], out["c_inv"][0], m[::3], 2, 0.01)
    out["knn_lab"], out["knn_co"], out["knn_di"] = np.packbits(lab), co, di
    R = oracle.Removerter(num_knn=2, knn_thr=0.01, schedule=[(0, 2.5), (1, 2.375), (0, 2.5)])
    for s, d in ((0, c), (1, q)):
        R.load_session(s, d.xyzi, d.offsets, d.poses, out["cq"[s] + "_inv"])
    R.run(step3=True)
    out["pipe_log"] = np.array([l[1:] for l in R.log()], np.int64)
    for n in ("nd_map", "pd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "updated_map"):
        out["pipe_" + n] = R.cloud("saved:" + n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "small_pair.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
