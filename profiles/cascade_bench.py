"""BASELINE configs[4]: S-session cascaded LT-map (default 6 sessions x 2000 keyframes) on 1..8 GPUs.  Session 0 is the live map; every
further session is diffed against it (Steps 0-3: Removerter.cpp:1656-1676) and the updated scans are promoted to the next live map
(ltrh_cascade_promote_updated: what the reference does through scans_updated/ + Session::loadKeyframes' VoxelGrid, Session.cpp:266-302).

  python profiles/cascade_bench.py [sessions] [keyframes]                                    1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/cascade_bench.py ...

With an even N the ranks are split by session (ranks < N/2 own the live map's keyframes, the others the query's); the per-stage times are
the max over ranks of host wall clock between stream synchronisations.  Rank 0 prints one JSON document."""
import json
import os
import sys
import time

import numpy as np

# torchrun exports OMP_NUM_THREADS=1 to its children; the host-side pieces here that use OpenMP (synthetic scan generation, the load-time
# VoxelGrid of a cascade promotion, the reference arm) want this rank's share of the cores.  Must happen before libgomp starts.
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _w = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    # session split: only the owners of the live session run the host-side VoxelGrid of a promotion, the other half of the ranks is idle then
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, _w // 2 if _w % 2 == 0 else _w)))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import synth  # noqa: E402
from lt_mapper_b200 import removert  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
SCHEDULE = [(0, 2.5)]          # the shipped run(): one removeOnce(2.5) per session (Removerter.cpp:1584, 1587)
NUM_KNN, KNN_THR = 2, 0.01     # params_ltmapper.yaml:65-66


def main():
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    split = world >= 2 and world % 2 == 0
    G = removert.Removerter(device=local, num_knn=NUM_KNN, knn_thr=KNN_THR, schedule=SCHEDULE)
    dist = None
    if world > 1:
        import torch.distributed as dist
        with bench._QuietStdout(to_stderr=True):      # NCCL prints its version banner on stdout
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            ids = [removert.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            G.init_nccl(ids[0], rank, world, split_sessions=split)
    own = bench.owned_blocks(rank, world, K, split)
    threads = max(1, (os.cpu_count() or 8) // max(1, world))
    empty = (np.zeros((0, 4), np.float32), np.zeros(1, np.int64), np.zeros((0, 4, 4)), np.zeros((0, 4, 4)))

    def block(slot, session):
        k0, n = own[slot]
        if n == 0:
            return empty
        d = synth.make_session(session, n, k0=k0, threads=threads)
        t = torch.from_numpy(d.xyzi).pin_memory()
        return (t.numpy(), d.offsets, d.poses, removert.inverse_poses(d.poses))

    def tmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def now():
        G.ctx.synchronize()
        return time.perf_counter()

    first = block(0, 0)
    nxt = block(1, 1)
    # warm-up (untimed): the first pair once through Steps 0-3 and a promotion, so that the timed stages do not pay for cudaMalloc,
    # page-locking of the staging buffers and NCCL channel set-up; then the live session is loaded afresh
    G.load_session(0, *first); G.load_session(1, *nxt)
    G.run_step0(); G.run_step12(); G.run_step3(); G.cascade_promote_updated()
    G.load_session(0, *first)
    stages = []
    if world > 1:
        dist.barrier()
    t_all = now()
    for s in range(1, S):
        t0 = now()
        G.load_session(1, *nxt)                     # pinned host -> device
        t1 = now()
        G.run_step0(); G.run_step12()
        t2 = now()
        G.run_step3()
        t3 = now()
        sizes = {}
        for n in ("nd_map", "pd_map", "strong_nd_map", "strong_pd_map", "updated_map"):
            try:
                sizes[n] = int(G.cloud_size("saved:" + n))
            except Exception:  # noqa: BLE001 -- held by the other group
                pass
        G.cascade_promote_updated()
        t4 = now()
        if s + 1 < S:
            nxt = block(1, s + 1)                   # generating the next session is not part of the measurement
        allsizes = [sizes]
        if world > 1:
            allsizes = [None] * world
            dist.all_gather_object(allsizes, sizes)
        merged = {}
        for d in allsizes:
            merged.update(d)
        st = {"query_session": s, "upload_ms": tmax(t1 - t0) * 1e3, "step0_12_ms": tmax(t2 - t1) * 1e3, "step3_ms": tmax(t3 - t2) * 1e3,
              "promote_ms": tmax(t4 - t3) * 1e3}
        st["keyframes_per_s"] = 2 * K / ((st["upload_ms"] + st["step0_12_ms"] + st["step3_ms"] + st["promote_ms"]) * 1e-3)
        st["promote_share"] = st["promote_ms"] / (st["upload_ms"] + st["step0_12_ms"] + st["step3_ms"] + st["promote_ms"])
        st.update(merged)
        stages.append(st)
        if world > 1:
            dist.barrier()
    total = sum(st["upload_ms"] + st["step0_12_ms"] + st["step3_ms"] + st["promote_ms"] for st in stages) * 1e-3
    if rank == 0:
        print(json.dumps({"workload": f"configs[4]: {S}-session cascade, {K} keyframes/session, 64x1800 scans, shipped schedule removeOnce(2.5), kNN k={NUM_KNN} thr={KNN_THR}",
                          "n_gpus": world, "parallelism": "session split + keyframe blocks" if split else "keyframe blocks",
                          "total_s": total, "keyframes_per_s_overall": 2 * K * (S - 1) / total, "stages": stages}, indent=1))
    G.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
