import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import synth, oracle
from lt_mapper_b200 import removert
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t = time.time(); c, q = synth.make_pair(K); print("synth", time.time() - t, c.xyzi.shape, flush=True)
sched = [(0, 2.5), (0, 2.0), (0, 1.5), (1, 1.0)]
G = removert.Removerter(num_knn=1, knn_thr=0.04, schedule=sched)
for s, d in ((0, c), (1, q)):
    G.load_session(s, d.xyzi, d.offsets, d.poses, oracle.inverse_poses(d.poses))
t = time.time(); G.run_step0(); print("step0", time.time() - t, G.cloud_size("map_global_curr_", 0), G.cloud_size("map_global_curr_", 1), flush=True)
for _ in range(2):
    G.run_step12(); G.reset_to_step0()
import os
if os.environ.get("LTR_TRACE") == "1": G.ctx.trace_dump(True)
t = time.time(); G.run_step12(); dt = time.time() - t
if os.environ.get("LTR_TRACE") == "1": G.ctx.trace_dump(True)
print("step12", dt, "kf/s", 2 * K / dt)
for k in ["hd_remove", "hd_knn", "parse_static", "ld_knn", "ld_filter", "ld_merge_viz"]:
    print(k, G.timing(k))
for l in G.log(): print(l)
print("launches", G.ctx.kernel_launches())
