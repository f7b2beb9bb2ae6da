"""Generates full-size golden digests by RUNNING THE REFERENCE (the ltremovert sources compiled unmodified from /root/reference
behind the third-party stand-ins, deterministic non-OpenMP build oracle/_ref/libltremovert_ref.so) on the synthetic pairs of
BASELINE.json:

  ref_config1_200kf.json   configs[1]: 200-keyframe pair, 64x1800 scans, remove@2.5/2.0/1.5 + revert@1.0, kNN k=1 thr 0.04
  ref_config2_100kf.json   configs[2] shape (selfRemovert [2.5, 2.0, 1.5], kNN k=2 thr 0.01, strong/weak split) at 100 keyframes --
                           the 1000-keyframe pair itself is hours of single-thread CPU time and is covered by the size-independent
                           properties and the N=1 vs N>1 digests of bench.py instead

Steps 0-2 (Removerter.cpp:1656-1669).  For every map the node holds and every PCD it saves afterwards: point count + SHA-256 of
the raw float32 bytes; for every per-keyframe scan set: the keyframe point counts' digest + one SHA-256 over all keyframes.
Needs /root/reference (make -C oracle ref).  Run from the repo root:  python tests/golden/make_ref_golden_fullsize.py [config1|config2]
Single-threaded on purpose (the reference's OpenMP per-pixel minimum is racy): ~20-40 min each."""
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from oracle import ref  # noqa: E402

MAPS = ["map_global_orig_", "map_global_curr_", "map_global_curr_static_", "map_global_curr_dynamic_", "map_global_nd_", "map_global_nd_strong_",
        "map_global_nd_weak_", "map_global_pd_", "map_global_pd_orig_", "map_global_pd_strong_", "map_global_pd_weak_"]
SCANSETS = ["keyframe_scans_", "keyframe_scans_static_projected_", "keyframe_scans_dynamic_", "scans_knn_coexist_", "scans_knn_diff_"]

CONFIGS = {
    "config1": dict(K=200, schedule=[(0, 2.5), (0, 2.0), (0, 1.5), (1, 1.0)], num_knn=1, knn_thr=0.04, out="ref_config1_200kf.json"),
    "config2": dict(K=100, schedule=None, selfremovert=[2.5, 2.0, 1.5], num_knn=2, knn_thr=0.01, out="ref_config2_100kf.json"),
}


def selfremovert_schedule(resolutions):
    s = []
    for r in resolutions:  # Removerter.cpp:1378-1393: remove(r), revert(0.95 r), remove(r); 0.95 * r is a double product narrowed once
        s += [(0, r), (1, float(np.float32(0.95 * float(np.float32(r))))), (0, r)]
    return s


def digest(a):
    a = np.ascontiguousarray(a)
    return [int(a.shape[0]), hashlib.sha256(a.tobytes()).hexdigest()]


def digest_scans(scans):
    h = hashlib.sha256()
    counts = np.array([len(a) for a in scans], np.int64)
    for a in scans:
        h.update(np.ascontiguousarray(a).tobytes())
    return [int(counts.sum()), hashlib.sha256(counts.tobytes()).hexdigest(), h.hexdigest()]


def run(name):
    cfg = CONFIGS[name]
    schedule = cfg["schedule"] or selfremovert_schedule(cfg["selfremovert"])
    c, q = synth.make_pair(cfg["K"])
    out = {"params": dict(K=cfg["K"], schedule=[[int(o), float(r)] for o, r in schedule], num_knn=cfg["num_knn"], knn_thr=cfg["knn_thr"],
                          seed=synth.SEED, beams=64, az_steps=1800)}
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        params = dict(save_pcd_directory=tmp + "/out/", sequence_vfov=50.0, sequence_hfov=360.0, ExtrinsicLiDARtoPoseBase=np.eye(4).ravel().tolist(),
                      downsample_voxel_size=0.05, num_nn_points_within=cfg["num_knn"], dist_nn_points_within=cfg["knn_thr"], num_omp_cores=1,
                      saveMapPCD=True)
        R = ref.Removerter(params, transform_order=0, write_files=False)
        for s, d in ((0, c), (1, q)):
            R.load_session_mem(s, d.xyzi, d.offsets, d.poses)
        R.stage("precleaningKeyframes"); R.stage("makeGlobalMap")
        print(name, "step0 done", round(time.time() - t0), "s", flush=True)
        R.high_dyn_with_schedule(schedule)
        print(name, "step1 done", round(time.time() - t0), "s", flush=True)
        R.stage("parseStaticScansViaProjection")
        R.stage("detectLowDynamicPoints")
        print(name, "step2 done", round(time.time() - t0), "s", flush=True)
        for s in (0, 1):
            for n in MAPS:
                out[f"map{s}:{n}"] = digest(R.cloud(n, s))
            for n in SCANSETS:
                out[f"scans{s}:{n}"] = digest_scans(R.scans(n, s))
        for path, a in R.saved():
            out["saved:" + os.path.relpath(path, tmp + "/out")] = digest(a)
        R.close()
    out["generated_in_s"] = round(time.time() - t0)
    with open(os.path.join(ROOT, "tests", "golden", cfg["out"]), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(name, len(out), "entries", out["generated_in_s"], "s")


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["config1", "config2"]):
        run(n)
