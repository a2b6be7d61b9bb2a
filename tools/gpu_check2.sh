#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -vv > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest.log | tail -30
timeout 600 python tools/diag_layers.py "((4,4),(3,12),(2,24))" 2 28 9 102 > gpurun_out/diag_three.log 2>&1; tail -20 gpurun_out/diag_three.log
timeout 600 python tools/diag_layers.py "((4,4),(3,6))" 2 28 8 101 > gpurun_out/diag_cfg2.log 2>&1; tail -12 gpurun_out/diag_cfg2.log
for wl in cfg1 one_eps cifar_2_6__2_24 three_eps_32; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_$wl.err
done
