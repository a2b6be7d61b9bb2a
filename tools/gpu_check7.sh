#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
R1=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_r1.so
( echo "== r1"; DCTN_B200_LIB=$R1 timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 5
  echo "== new"; timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 5
  for bn in 3:128 3:160 3:192 1:128 1:160 1:192; do echo "== new FAST_BN=$bn"; DCTN_B200_FAST_BN=$bn timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,input --train --iters 5; done
  echo "== r1 cifar"; DCTN_B200_LIB=$R1 timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 3
  echo "== new cifar"; timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 3
) > gpurun_out/kbench_ab.log 2>&1
grep -v "^$" gpurun_out/kbench_ab.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
for wl in cfg5_chain cfg5_convsbs; do timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err; cut -c1-600 gpurun_out/bench_$wl.json; done
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg5_chain --graph off --no-cpu-baseline > gpurun_out/bench_cfg5_chain_nograph.json 2>/dev/null; cut -c1-300 gpurun_out/bench_cfg5_chain_nograph.json
