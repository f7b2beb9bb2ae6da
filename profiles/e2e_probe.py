"""Phase timing of the end-to-end arm (pinned host -> H2D -> Step 0 -> step -> D2H)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lt_mapper_b200 import removert
K = int(sys.argv[1]) if len(sys.argv) > 1 else bench.KF_PER_SESSION
blocks = bench.gen_block(0, K)
pinned = [(torch.from_numpy(d.xyzi).pin_memory(), d.offsets, d.poses, inv) for d, inv in blocks]
R = removert.Removerter(num_knn=bench.NUM_KNN, knn_thr=bench.KNN_THR, schedule=bench.schedule())
def sync(): R.ctx.synchronize()
for it in range(4):
    t0 = time.perf_counter()
    for s, (t, o, p, ip) in enumerate(pinned): R.load_session(s, t.numpy(), o, p, ip)
    sync(); t1 = time.perf_counter()
    R.run_step0(); sync(); t2 = time.perf_counter()
    R.run_step12(); sync(); t3 = time.perf_counter()
    n = sum(R.cloud("saved:" + nm).nbytes for nm in bench.LD_OUTPUTS); t4 = time.perf_counter()
    print(f"it{it}: upload {1e3*(t1-t0):.1f} ms  step0 {1e3*(t2-t1):.1f}  step12 {1e3*(t3-t2):.1f}  download {1e3*(t4-t3):.1f} ({n/1e6:.0f} MB)  total {1e3*(t4-t0):.1f}")
    if it == 2 and os.environ.get("LTR_TRACE") == "1": R.ctx.trace_dump(True)
if os.environ.get("LTR_TRACE") == "1": R.ctx.trace_dump(True)
