"""Target for `ncu --set full`: a few launches of the dominant kernel (map projection of an HD remove pass and of the
visible-point extraction) on the configs[1] map.  Numbers printed under ncu are never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lt_mapper_b200 as ltr
from lt_mapper_b200 import removert

kf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
blocks = bench.gen_block(0, kf)
R = removert.Removerter(num_knn=1, knn_thr=0.04, schedule=bench.SCHEDULE)
for s, (d, inv) in enumerate(blocks):
    R.load_session(s, d.xyzi, d.offsets, d.poses, inv)
R.run_step0()
ctx = R.ctx
mh = R.cloud_handle("map_global_curr_", 0)
ss = R.scanset_handle("keyframe_scans_", 0)
n = ctx.remove_pass(mh, ss, 0, ltr.MODE_HD, 2.5)
vis = ctx.parse_projected(mh, 0, 0, kf, 3.0)
print("N", ctx.cloud_size(mh), "flagged", n, ctx.last_pass_stats())
