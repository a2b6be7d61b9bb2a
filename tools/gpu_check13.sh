#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( for db in 0 1; do echo "== DBUF=$db"; DCTN_B200_FAST_DBUF=$db timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,input --train --iters 2; done
) > gpurun_out/kbench_timing6.log 2>&1
grep -v "^$" gpurun_out/kbench_timing6.log | awk '!seen[$0]++' | cut -c1-420 | tail -40
