"""Turns the raw artefacts of profiles/scripts/r02_final_single_gpu.sh (under gpurun_out/) into the committed round-2 summaries:
  profiles/r02_launches.{csv,md}            ncu launch list of one bench step (per-kernel shares)
  profiles/r02_ncu_map_project_fast.md      `ncu --set full` of the dominant kernel: utilisation, stalls, traffic + source-level counters
  profiles/r02_traffic.json                 DRAM traffic per launch vs algorithmic bytes (read by bench.py for roofline.traffic)
  profiles/r02_sass_map_project_fast.md     static SASS mix + `ptxas -v` of the hot loop
Run here (no GPU needed; `ncu -i` only reads the report):  python profiles/make_r02_summaries.py"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def ncu_csv(rep, page, extra=()):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return txt


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    return [{h: r[ix[h]] for h in hdr} for r in rows[2:]], {h: units[ix[h]] for h in hdr}


def source_counters(rep):
    txt = ncu_csv(rep, "source", ("--print-source", "sass"))
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    out = []
    for b in blocks:
        hdr = b["rows"][0]
        ix = {h: i for i, h in enumerate(hdr)}
        ins = [r for r in b["rows"][1:] if len(r) > 10]
        out.append((b["name"], hdr, ix, ins))
    return out


def f(x):
    return float(str(x).replace(",", "")) if x not in ("", "n/a") else 0.0


def kernel_report(rep, title, n_map, kf_per_launch, lines):
    if not os.path.exists(rep):
        lines.append(f"\n## {title}\n\n(report {os.path.basename(rep)} missing)\n")
        return None
    mets, units = raw_metrics(rep)
    want = [("gpu__time_duration.sum", "kernel time"), ("launch__registers_per_thread", "registers / thread"), ("launch__occupancy_limit_registers", "CTAs / SM (register limit)"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64 / SM)"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
            ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (of 4)"), ("smsp__inst_executed.sum", "warp instructions"),
            ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
            ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate"), ("lts__t_sector_hit_rate.pct", "L2 sector hit rate"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)")]
    lines.append(f"\n## {title}\n")
    lines.append("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(mets))) + " |\n|---|" + "---|" * len(mets))
    for key, name in want:
        if key in mets[0]:
            lines.append(f"| {name} [{units[key]}] | " + " | ".join(f"{f(m[key]):,.2f}" for m in mets) + " |")
    traffic = None
    if "dram__bytes_read.sum" in mets[0]:
        def to_bytes(m, k):
            v = f(m[k]); u = units[k].lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
        per = [to_bytes(m, "dram__bytes_read.sum") + to_bytes(m, "dram__bytes_write.sum") for m in mets]
        traffic = sum(per) / len(per)
        alg = kf_per_launch * (12.0 * n_map + n_map / 8.0)
        lines.append(f"\nDRAM traffic per launch {traffic / 1e6:.0f} MB vs {alg / 1e6:.0f} MB algorithmic ({kf_per_launch} keyframes x (12 N + N/8), N = {n_map}): ratio {traffic / alg:.3f} -- the map tile is read once "
                     f"per {kf_per_launch} keyframes and the images stay in L2; the kernel is bound by instruction issue, not by HBM.")
    # source-level counters of the first kernel in the report
    src = source_counters(rep)
    if src:
        name, hdr, ix, ins = src[0]
        tot = sum(int(r[ix["Instructions Executed"]]) for r in ins)
        pairs32 = n_map * kf_per_launch / 32.0
        cnt = collections.Counter(int(r[ix["Instructions Executed"]]) for r in ins)
        hot_n, hot_c = max(cnt.items(), key=lambda kv: kv[0] * kv[1])
        hot = [r[ix["Source"]].strip() for r in ins if int(r[ix["Instructions Executed"]]) == hot_n]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for s in hot)
        stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        samples = sum(int(r[ix["Warp Stall Sampling (All Samples)"]]) for r in ins)
        stalls = {h[6:]: sum(int(r[ix[h]]) for r in ins) for h in stall_names}
        lines.append(f"\nSource-level counters (launch 1): {tot:,} warp instructions = **{tot / pairs32:.1f} per 32 (point, keyframe) pairs** (culled pairs included in the denominator). "
                     f"The keyframe step of a warp is {hot_c} SASS instructions executed {hot_n:,} times each ({hot_c / 4:.1f} per 32 evaluated pairs): " +
                     ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + ".")
        lines.append("\nWarp-stall samples (all): " + ", ".join(f"{k} {100.0 * v / max(samples, 1):.1f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]) + ".")
    return traffic


def main():
    # launch list
    src = os.path.join(OUT, "r02_launches.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "r02_launches.csv"))
        md = subprocess.run([sys.executable, os.path.join(PROF, "summarize_launches.py"), src, "40"], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, "r02_launches.md"), "w") as fo:
            fo.write("# Round 2: ncu launch list of ONE bench step (BASELINE configs[2]: 1000-keyframe pair, full selfRemovert)\n\n"
                     "Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv "
                     "python profiles/ncu_target.py` (1x B200, part of `profiles/scripts/r02_final_single_gpu.sh`). Cold-cache, serialised: compare shares, not absolutes. "
                     "No library kernel (CUB, cuBLAS, ...) appears: every launch is one of this repository's kernels.\n\n" + md)
    # ncu full
    n_map, kfb = 6864501, 32
    for line in open(os.path.join(OUT, "r02_ncu_hd.log"), errors="ignore") if os.path.exists(os.path.join(OUT, "r02_ncu_hd.log")) else []:
        m = re.match(r"N (\d+) flagged", line)
        if m:
            n_map = int(m.group(1))
    lines = ["# Round 2: `ncu --set full --clock-control none --import-source on` of the dominant kernel\n",
             f"Target: `profiles/ncu_kernel_target.py 200` (first HD remove pass and the visible-point extraction on the 200-keyframe map, N = {n_map} points, {kfb} keyframes per launch). "
             "Numbers under ncu are never bench values; they explain them."]
    t = kernel_report(os.path.join(OUT, "r02_fast_hd.ncu-rep"), "map_project_fast_kernel<candidates = true> (remove / revert / PD passes)", n_map, kfb, lines)
    kernel_report(os.path.join(OUT, "r02_fast_deferred.ncu-rep"), "map_project_fast_kernel<candidates = false, deferred = true> (ND pass, visible-point extraction)", n_map, kfb, lines)
    with open(os.path.join(PROF, "r02_ncu_map_project_fast.md"), "w") as fo:
        fo.write("\n".join(lines) + "\n")
    if t:
        json.dump({"N": n_map, "keyframes_per_launch": kfb, "dram_bytes_per_launch": t, "algorithmic_bytes_per_launch": kfb * (12.0 * n_map + n_map / 8.0),
                   "source": "profiles/r02_ncu_map_project_fast.md"}, open(os.path.join(PROF, "r02_traffic.json"), "w"), indent=1)
    # static SASS mix
    md = subprocess.run([sys.executable, os.path.join(PROF, "sass_mix.py")], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, "r02_sass_map_project_fast.md"), "w") as fo:
        fo.write(md + "\nDynamic counts (ncu source counters) are in `profiles/r02_ncu_map_project_fast.md`.\n")
    for name in ("r02_bench.json", "r02_bench_reference.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json", "r02_cascade_n8.json", "r02_pytest_gpu.txt", "r02_smoke.txt",
                 "r02_sanitizer.txt"):
        if os.path.exists(os.path.join(OUT, name)):
            shutil.copy(os.path.join(OUT, name), os.path.join(PROF, name))
    print("done")


if __name__ == "__main__":
    main()
