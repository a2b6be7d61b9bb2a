#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -25
timeout 600 python tools/bias_probe.py > gpurun_out/bias_probe2.log 2>&1; cat gpurun_out/bias_probe2.log | tail -14
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_c.json 2> gpurun_out/bench_cfg2_c.err; tail -2 gpurun_out/bench_cfg2_c.err
timeout 600 python benchmarks/logmatmulexp_bench.py --sizes 64,128,256,300 --iters 50 --json gpurun_out/lme_bench.json > gpurun_out/lme_bench.log 2>&1; tail -12 gpurun_out/lme_bench.log
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( for ord in 0 1 2; do for nt in 32 64 96; do echo "== dcore skip-gen order=$ord NT=$nt"; DCTN_B200_MMA_ORDER=$ord DCTN_B200_SKIP_GEN=1 DCTN_B200_DCORE_NT=$nt timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2; done; done
  echo "== full"; timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 3
) > gpurun_out/kbench_timing2.log 2>&1
grep -v "^$" gpurun_out/kbench_timing2.log | awk '!seen[$0]++' | cut -c1-420 | tail -40
