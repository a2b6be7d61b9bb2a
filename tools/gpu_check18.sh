#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( for m in 2 3; do echo "== SPLIT_M=$m"; DCTN_B200_SPLIT_M=$m timeout 300 python tools/kbench.py --layers c12 --batch 64 --kinds fwd,core,input --train --iters 2; done
  echo "== c23 m=2";  timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 2 ) > gpurun_out/kbench_timing7.log 2>&1
grep -v "^$" gpurun_out/kbench_timing7.log | awk '!seen[$0]++' | cut -c1-420 | tail -60
