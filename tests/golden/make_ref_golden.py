"""Generates tests/golden/ref_small_pair.json by RUNNING THE REFERENCE: the ltremovert sources compiled unmodified from
/root/reference behind the third-party stand-ins (oracle/_ref/libltremovert_ref.so, see oracle/ref_shim/include/ltr_shim_core.h)
execute Removerter::run()'s stages on the inputs stored in small_pair.npz; for every cloud the node holds or saves afterwards
the point count and the SHA-256 of its raw float32 bytes are stored (the comparisons are bit-exact, so digests lose nothing and
keep the fixture small), together with the inverse poses the reference computed.  Needs /root/reference (build: make -C oracle ref).  Run from the repo root:  python tests/golden/make_ref_golden.py

The fixture lets the oracle (CPU) and the CUDA path (GPU box, where neither /root/reference nor a compiler is needed) be
checked against outputs of the reference's own code."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

STAGES = ["precleaningKeyframes", "makeGlobalMap", "removeHighDynamicPoints", "parseStaticScansViaProjection", "detectLowDynamicPoints",
          "updateCurrentMap", "parseUpdatedStaticScansViaProjection", "parseLDScansViaProjection", "updateScansScanwise", "saveAllTypeOfScans"]
MAPS = ["map_global_orig_", "map_global_curr_", "map_global_curr_static_", "map_global_curr_dynamic_", "map_global_nd_", "map_global_nd_strong_",
        "map_global_nd_weak_", "map_global_pd_", "map_global_pd_orig_", "map_global_pd_strong_", "map_global_pd_weak_", "map_global_updated_",
        "map_global_updated_strong_"]
SCANSETS = ["keyframe_scans_", "keyframe_scans_static_projected_", "keyframe_scans_dynamic_", "scans_knn_coexist_", "scans_knn_diff_",
            "keyframe_scans_updated_", "keyframe_scans_updated_strong_", "keyframe_scans_pd_", "keyframe_scans_strong_pd_", "keyframe_scans_strong_nd_",
            "keyframe_scans_weak_nd_"]
PARAMS = dict(num_knn=2, knn_thr=0.01, voxel=0.05, order=0)


def digest(a):
    a = np.ascontiguousarray(a)
    return [int(a.shape[0]), hashlib.sha256(a.tobytes()).hexdigest()]


def main():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "small_pair.npz")))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        params = dict(save_pcd_directory=tmp + "/out/", sequence_vfov=50.0, sequence_hfov=360.0, ExtrinsicLiDARtoPoseBase=np.eye(4).ravel().tolist(),
                      downsample_voxel_size=PARAMS["voxel"], num_nn_points_within=PARAMS["num_knn"], dist_nn_points_within=PARAMS["knn_thr"],
                      num_omp_cores=1, saveMapPCD=True)
        R = ref.Removerter(params, transform_order=PARAMS["order"], write_files=False)
        R.load_session_mem(0, g["c_xyzi"], g["c_off"], g["c_poses"])
        R.load_session_mem(1, g["q_xyzi"], g["q_off"], g["q_poses"])
        for s in (0, 1):
            out[f"inv{s}"] = [float(v) for v in R.keyframe_poses(s)[1].ravel()]
        for st in STAGES:
            R.stage(st)
        for s in (0, 1):
            for n in MAPS:
                out[f"map{s}:{n}"] = digest(R.cloud(n, s))
            for n in SCANSETS:
                sc = R.scans(n, s)
                out[f"scans{s}:{n}"] = [digest(a) for a in sc]
        for path, a in R.saved():
            rel = os.path.relpath(path, tmp + "/out")
            out["saved:" + rel] = digest(a)
        R.close()
    out["params"] = PARAMS
    with open(os.path.join(ROOT, "tests", "golden", "ref_small_pair.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(len(out), "entries")


if __name__ == "__main__":
    main()
