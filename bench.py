#!/usr/bin/env python
"""bench.py -- keyframes/sec of the LT-removert + two-session diff hot path on B200 (BASELINE.json metric).

One "step" = Step 1 + static projection + Step 2 of Removerter::run() (ltremovert/src/Removerter.cpp:1665-1669:
removeHighDynamicPoints, parseStaticScansViaProjection, detectLowDynamicPoints) over one synthetic two-session pair.
Workload (BASELINE.json configs[1]): 200 keyframes per session PER GPU, 64 beams x 1800 steps (~103 k returns/scan),
3 remove + 1 revert resolutions, kNN diff k = 1, r = 0.2 m (threshold r^2 = 0.04 on the squared distance).

  python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference [...]                      the reference's CPU path (oracle, reference threading)

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definitions of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KF_PER_GPU = 200                                   # keyframes per session per GPU (configs[1])
SCHEDULE = [(0, 2.5), (0, 2.0), (0, 1.5), (1, 1.0)]  # 3 remove + 1 revert resolutions
NUM_KNN, KNN_THR = 1, 0.04                         # k = 1, r = 0.2 m  ->  r^2 on the squared distance (Session.cpp:592-596)
CPU_SAMPLE_KF = 20                                 # keyframes per session of the bounded CPU sample (~10-20 s of CPU work per step)
METRIC = "keyframes/sec (two-session removert+diff)"
LD_OUTPUTS = ["nd_map", "pd_map", "strong_nd_map", "weak_nd_map", "strong_pd_map", "weak_pd_map", "union_map_queryside",
              "union_map_centralside"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs (a background thread
    calling nvmlDeviceGetClockInfo / nvmlDeviceGetCurrentClocksEventReasons; `nvidia-smi -lms` was measured to slow the
    launch-heavy step by ~30 % through driver-lock contention, direct NVML calls do not)."""

    def __init__(self, gpu_index, period_s=0.05):
        self.gpu, self.period, self.rows, self.stop_flag, self.thread, self.err = gpu_index, period_s, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        self.stop_flag = True
        self.thread.join(timeout=5)
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        reasons = set()
        for _, rs in self.rows:
            for n, bit in names.items():
                if rs & bit:
                    reasons.add(n)
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml"}


def gen_block(rank, kf):
    import synth
    import oracle  # inverse poses only (4x4 double inverse; the C-ABI takes inverse poses as an input)
    threads = max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))
    out = []
    for s in (0, 1):
        d = synth.make_session(s, kf, k0=rank * kf, threads=threads)
        out.append((d, np.stack([np.linalg.inv(p) for p in d.poses])))
    del oracle
    return out


class _QuietStdout:
    """Keeps bench.py's stdout to the one JSON line: native code that narrates on fd 1 (the reference on std::cout, Session.cpp:575;
    NCCL's version banner) is sent to /dev/null or, with to_stderr, to fd 2 for the duration."""

    def __init__(self, to_stderr=False):
        self._to_stderr = to_stderr

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._dst = os.dup(2) if self._to_stderr else os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._dst, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._saved); os.close(self._dst)


def cpu_baseline_run(steps=1, warmup=0):
    """The reference's CPU path on a bounded sample of the SAME workload: the first CPU_SAMPLE_KF keyframes of each session,
    same schedule and kNN parameters, Step 1 + static projection + Step 2.

    kind "reference": the reference's own ltremovert sources, compiled unmodified behind the third-party stand-ins
    (oracle/_ref/libltremovert_ref_omp.so, OpenMP pragmas active as in the reference's build; see oracle/ref_shim), driven
    through its own member functions.  kind "port" (only when that library was not built): the oracle in `faithful` mode
    (reference threading structure).  Threads: num_omp_cores = min(16, host cores); map2RangeImg hard-codes 16 (utility.cpp:109)."""
    import synth
    c, q = synth.make_pair(CPU_SAMPLE_KF)
    cores = min(16, os.cpu_count() or 1)
    from oracle import ref
    use_ref = ref.available(omp=True)
    times = []
    for it in range(warmup + steps):
        if use_ref:
            import tempfile
            with tempfile.TemporaryDirectory() as tmp, _QuietStdout():
                R = ref.Removerter(dict(save_pcd_directory=tmp + "/", sequence_vfov=50.0, sequence_hfov=360.0,
                                        ExtrinsicLiDARtoPoseBase=np.eye(4).ravel().tolist(), downsample_voxel_size=0.05,
                                        num_nn_points_within=NUM_KNN, dist_nn_points_within=KNN_THR, num_omp_cores=cores),
                                   omp=True, write_files=False)
                for s, d in ((0, c), (1, q)):
                    R.load_session_mem(s, d.xyzi, d.offsets, d.poses)
                R.stage("precleaningKeyframes"); R.stage("makeGlobalMap")
                t0 = time.perf_counter()
                R.high_dyn_with_schedule(SCHEDULE)
                R.stage("parseStaticScansViaProjection")
                R.stage("detectLowDynamicPoints")
                dt = time.perf_counter() - t0
                n_map = len(R.cloud("map_global_orig_", 0))
                R.close()
        else:
            import oracle
            R = oracle.Removerter(num_knn=NUM_KNN, knn_thr=KNN_THR, schedule=SCHEDULE, faithful=True, omp_cores=cores, threads=cores)
            for s, d in ((0, c), (1, q)):
                R.load_session(s, d.xyzi, d.offsets, d.poses, np.stack([np.linalg.inv(p) for p in d.poses]))
            R.run(step0=True, step12=False)
            t0 = time.perf_counter()
            R.run(step0=False, step12=True)
            dt = time.perf_counter() - t0
            n_map = len(R.cloud("map_global_orig_", 0))
            del R
        if it >= warmup:
            times.append(dt)
    dt = float(np.mean(times))
    how = ("the reference's own ltremovert sources compiled behind third-party stand-ins (oracle/_ref, OpenMP on; PCL/FLANN/Eigen calls go to the "
           "oracle's restatements)" if use_ref else "oracle in reference-threading mode (oracle/_ref not built)")
    return {"value": 2 * CPU_SAMPLE_KF / dt, "unit": "keyframes/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"first {CPU_SAMPLE_KF} keyframes of each session of the same synthetic pair (64x1800 scans, {n_map} merged points), "
                      f"same schedule/kNN, Step 1 + static projection + Step 2, {how}; {dt:.2f} s per step. "
                      f"Cost is O(K*N): per-keyframe CPU cost at the full 200-keyframe map is higher, so this ratio is conservative",
            "seconds_per_step": dt}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = cpu_baseline_run(steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "keyframes/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (f64 transforms)", "data": "synthetic",
            "config": {"workload": "configs[1] schedule on a bounded CPU sample (see cpu_baseline.sample)", "schedule": SCHEDULE,
                       "num_knn": NUM_KNN, "knn_thr": KNN_THR, "keyframes_per_session": CPU_SAMPLE_KF},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kf", type=int, default=KF_PER_GPU, help="keyframes per session per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-path", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: measure the sampler's own overhead")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import torch
    import torch.distributed as dist
    from lt_mapper_b200 import removert

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lt_mapper_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        with _QuietStdout(to_stderr=True):     # NCCL prints its version banner on stdout at communicator creation
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        comm = removert.TorchDistComm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blocks = gen_block(rank, args.kf)
    # pinned host copies: the e2e arm's H2D source
    pinned = []
    for d, inv in blocks:
        t = torch.from_numpy(d.xyzi).pin_memory()
        pinned.append((t, d.offsets, d.poses, inv))
    h2d_bytes = sum(int(t.numel()) * 4 + o.nbytes + p.nbytes + ip.nbytes for t, o, p, ip in pinned)

    R = removert.Removerter(device=local_rank, num_knn=NUM_KNN, knn_thr=KNN_THR, schedule=SCHEDULE, comm=comm,
                            fast_path=not args.no_fast_path)

    def load():
        for s, (t, o, p, ip) in enumerate(pinned):
            R.load_session(s, t.numpy(), o, p, ip)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---------------- device-resident arm: `value` ----------------
    load()
    R.run_step0()
    n_map = [R.cloud_size("map_global_curr_", s) for s in (0, 1)]
    for _ in range(args.warmup):
        R.reset_to_step0(); R.run_step12()
    barrier()
    R.ctx.profile_reset()
    sampler = ClockSampler(local_rank)
    if not args.no_clock_sampler:
        sampler.start()
    l0 = R.ctx.kernel_launches()
    stage_t = {}
    barrier()
    R.ctx.timer_start()
    for _ in range(args.steps):
        flush.zero_()                 # L2 flush between timed iterations
        R.reset_to_step0()
        R.run_step12()
        for k in ("hd_remove", "hd_knn", "parse_static", "ld_knn", "ld_filter", "ld_merge_viz"):
            stage_t[k] = stage_t.get(k, 0.0) + R.timing(k)
    ms_total = R.ctx.timer_stop()
    barrier()
    if os.environ.get("LTR_ALLOC_STATS") == "1":
        R.ctx.trace_dump(False)
    launches = R.ctx.kernel_launches() - l0
    clocks = sampler.stop()
    prof = R.ctx.profile_get()
    passlog = R.log()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    total_kf = 2 * args.kf * world
    value = total_kf / (ms_step * 1e-3)

    # ---------------- end-to-end arm: host buffers -> host ND/PD maps ----------------
    pinned_out = {}

    def e2e_step():
        load()                        # H2D from pinned memory, inside the timed region
        R.run_step0()
        R.run_step12()
        n = 0
        for name in LD_OUTPUTS:
            try:
                h = R.cloud_handle("saved:" + name)
            except Exception:
                continue
            need = R.ctx.cloud_size(h)
            if name not in pinned_out or len(pinned_out[name]) < need:      # pinned destination, grown on demand (warm-up)
                pinned_out[name] = torch.empty((int(need * 1.25) + 1024, 4), dtype=torch.float32).pin_memory().numpy()
            n += R.ctx.cloud_download(h, out=pinned_out[name]).nbytes      # D2H of the merged ND/PD / union maps
        return n
    e2e_iter_ms = []
    for _ in range(max(3, args.warmup)):
        t1 = time.perf_counter(); e2e_step(); e2e_iter_ms.append(round((time.perf_counter() - t1) * 1e3, 1))
    barrier()
    t0 = time.perf_counter()
    R.ctx.timer_start()
    d2h_bytes = 0
    for _ in range(args.steps):
        flush.zero_()
        t1 = time.perf_counter(); d2h_bytes = e2e_step(); e2e_iter_ms.append(round((time.perf_counter() - t1) * 1e3, 1))
    ms_e2e_dev = R.ctx.timer_stop()
    barrier()
    ms_e2e = max(ms_e2e_dev, (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_kf / (float(t.item()) / args.steps * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel: map projection of the remove/revert passes
        k_us = prof[0] / max(prof[1], 1.0)
        k_bytes = prof[2] / max(prof[1], 1.0)
        achieved = (k_bytes / (k_us * 1e-6)) / 1e9 if k_us > 0 else 0.0
        traffic, traffic_note = None, "no ncu capture committed"
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"]
            traffic_note = (f"ncu --set full (profiles/r01_ncu_map_project_fast.md): dram__bytes_read+write per 32-keyframe launch on the N={tj['N']} map = "
                            f"{tj['dram_bytes_per_launch'] / 1e6:.0f} MB vs {tj['algorithmic_bytes_per_launch'] / 1e6:.0f} MB algorithmic "
                            "(the map tile is read once per 32 keyframes; images stay in L2) -> the kernel is instruction-issue bound, not DRAM bound")
        roofline = {"bound": "hbm", "kernel": "map_project_fast_kernel<candidates> (remove/revert/PD passes; ND launches included)", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                    "launches_timed": int(prof[1]), "avg_launch_us": k_us, "algorithmic_bytes_per_launch": k_bytes,
                    "point_projections_per_s": prof[3] / (prof[0] * 1e-6) if prof[0] > 0 else 0.0,
                    "kernel_share_of_step": (prof[0] / args.steps) / (ms_step * 1e3),
                    "parse_kernel": {"avg_launch_us": prof[4] / max(prof[5], 1.0), "achieved": (prof[6] / max(prof[4], 1e-9)) * 1e6 / 1e9}}
        line = {"metric": METRIC, "value": value, "unit": "keyframes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (f64 transforms)", "data": "synthetic",
                "config": {"workload": f"configs[1]: {args.kf}-keyframe two-session pair per GPU ({args.kf * world} keyframes/session total, "
                                       f"keyframe-sharded), 64x1800 scans, 3 remove + 1 revert resolutions, kNN k=1 r=0.2m",
                           "schedule": SCHEDULE, "num_knn": NUM_KNN, "knn_thr": KNN_THR, "points_per_scan": float(np.mean(np.diff(blocks[0][0].offsets))),
                           "map_points": n_map, "timed_region": "Step 1 + static projection + Step 2 (Removerter.cpp:1665-1669), inputs resident in HBM",
                           "l2": "256 MiB buffer written between timed iterations", "fast_path": not args.no_fast_path,
                           "parallelism": f"keyframe-sharded x{world}, maps replicated"},
                "clocks": clocks, "gpu_launches": int(launches / args.steps),
                "e2e": {"value": e2e_value, "unit": "keyframes/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "region": "pinned host scans+poses -> H2D -> Step 0 + Step 1 + static projection + Step 2 -> D2H of the ND/PD/union maps",
                        "iteration_ms_incl_warmup": e2e_iter_ms},
                "roofline": roofline,
                "stages_ms_per_step": {k: v / args.steps * 1e3 for k, v in stage_t.items()},
                "pass_log": passlog[:8]}
        if comm is not None:
            line["comm_hooks_rank0_total"] = {k: {"calls": v[0], "seconds": round(v[1], 4), "bytes": v[2]} for k, v in comm.stats.items()}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline_run()
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    R.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
