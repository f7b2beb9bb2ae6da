"""Writes tests/golden/config2_pass_sizes.json from the pass log of a device run of BASELINE configs[2] (a JSON line of bench.py or
of profiles/config2_probe.py given as argv[1]): the map size every pass of the step sees.  These sizes are properties of the workload
(seeded synthetic pair + schedule), not of the implementation; tests/test_gpu_fullsize.py::test_config2_pass_sizes re-derives them on the
GPU box and bench.py --impl reference checks the first of them against the map the reference itself builds."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
log = line["pass_log"]
hd = [e for e in log if e[0] in ("removeOnce", "revertOnce")]
half = len(hd) // 2
# the two sessions are told apart by their first pass (the full map); the central one has the smaller map in this pair
first = sorted([e for e in hd if e[0] == "removeOnce"], key=lambda e: -e[1])[:2]
chains = []
for start in sorted(first, key=lambda e: e[1]):
    i = hd.index(start)
    chains.append(hd[i:i + half])
out = {"keyframes_per_session": line.get("config", {}).get("keyframes_per_session", line.get("keyframes", 2000) // 2),
       "map_points": [c[0][1] for c in chains],
       "hd_pass_map_points": [[e[1] for e in c] for c in chains],
       "hd_pass_dynamic": [[e[2] for e in c] for c in chains],
       "static_map_points": [c[-1][3] for c in chains],
       "nd_pass_map_points": [e[1] for e in log if e[0] == "iremoveOnceForND"],
       "pd_pass_map_points": [e[1] for e in log if e[0] == "removeOnceForPD"],
       "source": "pass log of " + os.path.basename(sys.argv[1])}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config2_pass_sizes.json"), "w"), indent=1)
print(out)
