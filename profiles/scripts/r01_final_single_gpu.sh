set -x
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Knn diff\|display every\|^$\|Read a pointcloud\|\.pcd$" | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err
python profiles/cascade_probe.py 6 100 > gpurun_out/r01_cascade.json 2> gpurun_out/cascade.err; tail -c 300 gpurun_out/cascade.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_v6.csv python profiles/ncu_target.py > gpurun_out/ncu_target.log 2>&1; tail -2 gpurun_out/ncu_target.log
python -c "
import json
d=json.load(open('gpurun_out/r01_bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])
r=json.load(open('gpurun_out/r01_bench_reference.json')); print('ref arm', r['value'], r['ms_per_step'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
c=json.load(open('gpurun_out/r01_cascade.json')); print('cascade', c['keyframes_per_s_overall'], [round(s['promote_ms']) for s in c['stages']])
"
