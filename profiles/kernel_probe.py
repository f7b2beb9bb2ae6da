"""Micro-benchmark of the projection kernels on the configs[1] map: per-pass CUDA-event time, exact-path share, atomics."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lt_mapper_b200 as ltr
from lt_mapper_b200 import removert

kf = int(sys.argv[1]) if len(sys.argv) > 1 else 200
blocks = bench.gen_block(0, kf)
for fast, cull in ((True, True), (True, False), (False, False)):
    R = removert.Removerter(num_knn=1, knn_thr=0.04, schedule=bench.SCHEDULE, fast_path=fast, cull=cull)
    for s, (d, inv) in enumerate(blocks):
        R.load_session(s, d.xyzi, d.offsets, d.poses, inv)
    R.run_step0()
    ctx = R.ctx
    mh = R.cloud_handle("map_global_curr_", 0)
    ss = R.scanset_handle("keyframe_scans_", 0)
    import ctypes
    # poses handle of session 0 is 0 by construction (first upload)
    ps = 0
    N = ctx.cloud_size(mh)
    for mode, alpha, name in ((ltr.MODE_HD, 2.5, "HD 2.5"), (ltr.MODE_HD, 1.0, "HD 1.0"), (ltr.MODE_ND, 2.5, "ND 2.5")):
        for rep in range(2):
            ctx.profile_reset()
            n = ctx.remove_pass(mh, ss, ps, mode, alpha)
            st = ctx.last_pass_stats()
            pf = ctx.profile_get()
        A = kf * (12 * N + N / 8)
        print(f"fast={fast} cull={cull} {name}: N={N} flagged={n} pass {st[4]/1e3:.2f} ms (map kernels {pf[0]/1e3:.2f} ms in {int(pf[1])} launches)  {st[0]/st[4]*1e6/1e9:.1f} Gpair/s  alg {A/st[4]*1e6/1e9:.0f} GB/s  exact share {st[2]/st[0]:.4f} culled {st[6]/st[0]:.3f} atomics/pair {st[3]/st[0]:.5f}")
    for rep in range(2):
        t = time.time(); vis = ctx.parse_projected(mh, ps, 0, kf, 3.0); ctx.synchronize(); dt = time.time() - t
        st = ctx.last_pass_stats()
        ctx.scanset_free(vis)
    print(f"fast={fast} cull={cull} parse 3.0: {st[4]/1e3:.2f} ms ({dt*1e3:.1f} wall) {st[0]/st[4]*1e6/1e9:.1f} Gpair/s exact share {st[2]/st[0]:.4f} atomics/pair {st[3]/st[0]:.5f}")
    R.close()
