"""Worker of tests/test_gpu_multi.py: one rank of a torchrun job (or a single process) runs Steps 0-3 on a small synthetic pair and
writes the SHA-256 of every cloud / per-keyframe cloud it holds to <out>/rank<r>.json.  Keys carry GLOBAL keyframe indices, so the
files of all ranks of a run can be merged and compared with the single-process run."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MAPS = ["map_global_orig_", "map_global_curr_", "map_global_curr_static_", "map_global_curr_dynamic_", "map_global_nd_", "map_global_nd_strong_",
        "map_global_nd_weak_", "map_global_pd_", "map_global_pd_orig_", "map_global_pd_strong_", "map_global_pd_weak_", "map_global_updated_",
        "map_global_updated_strong_"]
SCANSETS = ["keyframe_scans_", "keyframe_scans_static_projected_", "keyframe_scans_dynamic_", "scans_knn_coexist_", "scans_knn_diff_",
            "keyframe_scans_updated_", "keyframe_scans_updated_strong_", "keyframe_scans_pd_", "keyframe_scans_strong_pd_", "keyframe_scans_strong_nd_",
            "keyframe_scans_weak_nd_"]
SAVED = ["OriginalNoisyCentralMapGlobal", "OriginalNoisyQueryMapGlobal", "central_sess_high_dyn", "query_sess_high_dyn", "union_map_queryside",
         "union_map_centralside", "pd_map", "nd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "updated_map", "updated_map_strong"]


def dg(a):
    a = np.ascontiguousarray(a)
    return [int(len(a)), hashlib.sha256(a.tobytes()).hexdigest()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--kf", type=int, default=8)
    ap.add_argument("--split", type=int, default=1)
    ap.add_argument("--cascade", type=int, default=0, help="after Step 3: promote + a second query session (same data) + Steps 0-3 again")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import bench
    import synth
    from lt_mapper_b200 import removert
    torch.cuda.set_device(local)
    split = bool(args.split) and world >= 2 and world % 2 == 0
    R = removert.Removerter(device=local, num_knn=2, knn_thr=0.01, schedule=removert.selfremovert_schedule([2.5, 2.0]))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [removert.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        R.init_nccl(ids[0], rank, world, split_sessions=split)
    own = bench.owned_blocks(rank, world, args.kf, split)
    empty = (np.zeros((0, 4), np.float32), np.zeros(1, np.int64), np.zeros((0, 4, 4)), np.zeros((0, 4, 4)))

    def load(sessions):
        for s in sessions:
            k0, n = own[s]
            if n == 0:
                R.load_session(s, *empty)
                continue
            d = synth.make_session(s, n, k0=k0, beams=16, az_steps=600, threads=2)
            R.load_session(s, d.xyzi, d.offsets, d.poses, removert.inverse_poses(d.poses))

    load((0, 1))
    R.run_step0(); R.run_step12(); R.run_step3()
    if args.cascade:
        R.cascade_promote_updated()
        load((1,))
        R.run_step0(); R.run_step12(); R.run_step3()
    out = {}
    for s in (0, 1):
        for n in MAPS:
            try:
                out[f"map{s}:{n}"] = dg(R.cloud(n, s))
            except Exception:  # noqa: BLE001 -- not held by this rank
                pass
        for n in SCANSETS:
            try:
                pts, off = R.scanset(n, s)
            except Exception:  # noqa: BLE001
                continue
            for k in range(len(off) - 1):
                out[f"scans{s}:{n}:{own[s][0] + k}"] = dg(pts[off[k]:off[k + 1]])
    for n in SAVED:
        try:
            out["saved:" + n] = dg(R.cloud("saved:" + n))
        except Exception:  # noqa: BLE001
            pass
    out["__log__"] = R.log()
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, f"rank{rank}.json"), "w") as f:
        json.dump(out, f)
    R.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
