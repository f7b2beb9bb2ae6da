"""torchrun target: keyframe-sharded run on WORLD_SIZE GPUs, checked on rank 0 against a single-GPU run of the same pair.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/multi_gpu_check.py [kf_per_gpu]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from lt_mapper_b200 import removert

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
kf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
beams, az = (16, 600) if kf <= 16 else (64, 1800)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sched = [(0, 2.5), (1, 2.375), (0, 2.5)]
NAMES = ["nd_map", "pd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "union_map_queryside", "union_map_centralside",
         "updated_map", "updated_map_strong", "central_sess_high_dyn"]

def run(comm, k0, nkf, device):
    R = removert.Removerter(device=device, num_knn=2, knn_thr=0.01, schedule=sched, comm=comm)
    for s in (0, 1):
        d = synth.make_session(s, nkf, beams=beams, az_steps=az, k0=k0, threads=8)
        R.load_session(s, d.xyzi, d.offsets, d.poses, np.stack([np.linalg.inv(p) for p in d.poses]))
    t = time.perf_counter(); R.run_step0(); R.run_step12(); R.run_step3(); R.ctx.synchronize(); dt = time.perf_counter() - t
    out = {n: R.cloud("saved:" + n) for n in NAMES}
    out["log"] = R.log()
    vis = R.scanset("keyframe_scans_updated_", 0)
    R.close()
    return out, vis, dt

multi, vis_m, dt_m = run(removert.TorchDistComm(), rank * kf, kf, local)
dist.barrier()
if rank == 0:
    single, vis_s, dt_s = run(None, 0, kf * world, local)
    ok = multi["log"] == single["log"]
    for n in NAMES:
        same = multi[n].shape == single[n].shape and np.array_equal(multi[n].view(np.uint32), single[n].view(np.uint32))
        ok &= same
        if not same: print("MISMATCH", n, multi[n].shape, single[n].shape)
    # rank 0's keyframe block of the per-keyframe outputs equals the first kf keyframes of the single run
    pm, om = vis_m; ps, os_ = vis_s
    same = np.array_equal(om, os_[:kf + 1]) and np.array_equal(pm.view(np.uint32), ps[:os_[kf]].view(np.uint32))
    ok &= same
    print(f"MULTI_GPU_CHECK world={world} kf/gpu={kf} identical={bool(ok)} t_multi={dt_m:.3f}s t_single={dt_s:.3f}s", flush=True)
dist.barrier()
dist.destroy_process_group()
