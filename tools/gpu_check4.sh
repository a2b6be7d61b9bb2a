#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest.log | tail -12
timeout 600 python tools/bias_probe.py > gpurun_out/bias_probe.log 2>&1; cat gpurun_out/bias_probe.log | tail -14
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_b.json 2> gpurun_out/bench_cfg2_b.err; tail -2 gpurun_out/bench_cfg2_b.err
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 3
  for nt in 64 96 128; do echo "== dcore NT=$nt"; DCTN_B200_DCORE_NT=$nt timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2; done
  echo "== dcore skip-gen"; DCTN_B200_SKIP_GEN=1 timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2
  for nt in 64 128; do echo "== dcore skip-gen NT=$nt"; DCTN_B200_SKIP_GEN=1 DCTN_B200_DCORE_NT=$nt timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2; done
) > gpurun_out/kbench_timing.log 2>&1
grep -v "^$" gpurun_out/kbench_timing.log | awk '!seen[$0]++' | cut -c1-420 | tail -60
