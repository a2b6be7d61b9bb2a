#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
for wl in cfg2 cifar_2_6__2_24 cifar_2_12__2_24 cifar_2_23__2_24 three_eps cfg1 one_eps; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_i.json 2>gpurun_out/bench_${wl}_i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_i.json')); print('$wl', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'])); [print('    ',k) for k in d['roofline']['all_kernels']]"; done
