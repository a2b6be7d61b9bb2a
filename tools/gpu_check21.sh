#!/bin/bash
# register-group generated operand in the generic GEMM: parity of the generic-path tests, cycle probes, timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "generic or cifar or three_layer or tc_forward_and_input or tch3 or saved" > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reg.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_reg.log | cut -c1-300 | tail -12
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( echo "== c12"; timeout 300 python tools/kbench.py --layers c12 --batch 64 --kinds fwd,core,input --train --iters 2
  echo "== c23";  timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 2
  echo "== k3q3";  DCTN_DEBUG_SHAPE=1 timeout 300 python tools/kbench.py --layers k3q3 --batch 512 --kinds fwd,core,input --train --iters 2 ) > gpurun_out/kbench_reg.log 2>&1
grep -v "^$" gpurun_out/kbench_reg.log | awk '!seen[$0]++' | cut -c1-420 | tail -40
unset DCTN_B200_LIB DCTN_TCG_DEBUG
for wl in cifar_2_12__2_24 cifar_2_23__2_24; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_reg.json 2>gpurun_out/bench_${wl}_reg.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_reg.json')); print('$wl', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'])); [print('    ',k) for k in d['roofline']['all_kernels']]"; done
