"""Target for `ncu`: one steady-state bench step (Step 1 + static projection + Step 2 on the configs[1] workload)
inside a cudaProfilerStart/Stop range.  Usage (on the GPU box):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/ncu_target.py
Numbers printed by a run under ncu are never bench values."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lt_mapper_b200 import removert  # noqa: E402

kf = int(sys.argv[1]) if len(sys.argv) > 1 else bench.KF_PER_GPU
blocks = bench.gen_block(0, kf)
R = removert.Removerter(num_knn=bench.NUM_KNN, knn_thr=bench.KNN_THR, schedule=bench.SCHEDULE)
for s, (d, inv) in enumerate(blocks):
    R.load_session(s, d.xyzi, d.offsets, d.poses, inv)
R.run_step0()
R.run_step12()
R.reset_to_step0()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
R.run_step12()
R.ctx.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches", R.ctx.kernel_launches(), {k: R.timing(k) for k in ("hd_remove", "hd_knn", "parse_static", "ld_knn", "ld_filter", "ld_merge_viz")})
