"""BASELINE configs[2]: 1x B200, 1000-keyframe pair, full selfRemovert (remove r, revert 0.95 r, remove r for r in 2.5, 2.0, 1.5)
+ strong/weak ND/PD split; one cold run + one timed warm run of Step 1 + static projection + Step 2 (and Step 3 once)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lt_mapper_b200 import removert

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.perf_counter(); blocks = bench.gen_block(0, K); t_gen = time.perf_counter() - t0
sched = removert.selfremovert_schedule([2.5, 2.0, 1.5])
R = removert.Removerter(num_knn=2, knn_thr=0.01, schedule=sched)
t0 = time.perf_counter()
for s, (d, inv) in enumerate(blocks):
    R.load_session(s, d.xyzi, d.offsets, d.poses, inv)
R.run_step0(); R.ctx.synchronize(); t_step0 = time.perf_counter() - t0
n_map = [R.cloud_size("map_global_curr_", s) for s in (0, 1)]
R.run_step12(); R.reset_to_step0()
R.ctx.profile_reset()
R.ctx.timer_start(); R.run_step12(); ms = R.ctx.timer_stop()
stages = {k: round(R.timing(k) * 1e3, 1) for k in ("hd_remove", "hd_knn", "parse_static", "ld_knn", "ld_filter", "ld_merge_viz")}
prof = R.ctx.profile_get()
log = R.log()
t0 = time.perf_counter(); R.run_step3(); R.ctx.synchronize(); t3 = time.perf_counter() - t0
out = {"workload": f"configs[2]: {K}-keyframe pair, selfRemovert [2.5, 2.0, 1.5], kNN k=2 thr 0.01, strong/weak ND/PD", "keyframes": 2 * K,
       "map_points": n_map, "synth_s": round(t_gen, 1), "upload_plus_step0_s": round(t_step0, 2), "step12_ms": round(ms, 1), "keyframes_per_s": round(2 * K / (ms * 1e-3), 1),
       "step3_s": round(t3, 2), "stages_ms": stages, "map_kernel": {"launches": int(prof[1]), "avg_us": round(prof[0] / max(prof[1], 1), 1),
       "algorithmic_GBps": round(prof[2] / max(prof[0], 1e-9) * 1e6 / 1e9, 1), "Gpairs_per_s": round(prof[3] / max(prof[0], 1e-9) * 1e6 / 1e9, 1)},
       "nd_pd": {n: R.cloud_size("saved:" + n) for n in ("nd_map", "pd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map")},
       "pass_log": log}
print(json.dumps(out))
