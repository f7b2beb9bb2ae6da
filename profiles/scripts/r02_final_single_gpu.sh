#!/bin/bash
# Round-2 single-GPU measurement sequence (one `gpurun` call from the repo root).  Everything lands in gpurun_out/; the summaries judged are
# copied / derived into profiles/ afterwards (profiles/README.md).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
# launch list of one steady-state step (cold-cache, serialised: compare shares, not absolutes)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python profiles/ncu_target.py > gpurun_out/r02_ncu_target.log 2>&1
# full capture of the dominant kernel: 3 launches of the scan-minus-map variant, 2 of the deferred true-minimum variant
ncu --set full --clock-control none --import-source on -k regex:map_project_fast -s 2 -c 3 -o gpurun_out/r02_fast_hd python profiles/ncu_kernel_target.py 200 > gpurun_out/r02_ncu_hd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_project_fast_kernel.*false -c 2 -o gpurun_out/r02_fast_deferred python profiles/ncu_kernel_target.py 200 > gpurun_out/r02_ncu_deferred.log 2>&1
# memory checker over the small pipeline tests (aux: SURVEY section 5)
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -q -x -k "shipped_run or multires" > gpurun_out/r02_sanitizer.txt 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r02_sanitizer.txt
tail -3 gpurun_out/r02_pytest_gpu.txt gpurun_out/r02_smoke.txt gpurun_out/r02_sanitizer.txt
