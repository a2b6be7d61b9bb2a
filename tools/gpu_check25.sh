#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -m gpu -q -x -k "logmatmulexp or lme or convsbs or conv_sbs" > gpurun_out/pytest_lme.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lme.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_lme.log | cut -c1-300 | tail -8
for wl in cfg5_chain cfg5_convsbs; do timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r02g_bench_$wl.json 2> gpurun_out/r02g_bench_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_$wl.json')); print('$wl', d['value'], d['unit'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'fwd call ms', d['roofline'].get('ms_per_call'))"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5_chain --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg5.log 2>&1
grep "lme" gpurun_out/launches_cfg5.csv | tail -12 | awk -F'","' '{print substr($5,1,60), $NF}'
