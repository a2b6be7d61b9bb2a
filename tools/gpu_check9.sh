#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( echo "== default"; timeout 300 python tools/kbench.py --layers L1,L2 --kinds core --iters 5
  for cfg in "2 96" "2 128" "1 128" "1 96" "0 128" "1 64" "3 96"; do set -- $cfg; echo "== BNH=$1 NT=$2"; DCTN_B200_DCORE_BNH=$1 DCTN_B200_DCORE_NT=$2 timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 5; done
  for cfg in "4 64" "3 128" "2 128" "2 64" "1 128" "0 128"; do set -- $cfg; echo "== L1 BNH=$1 NT=$2"; DCTN_B200_DCORE_BNH=$1 DCTN_B200_DCORE_NT=$2 timeout 300 python tools/kbench.py --layers L1 --kinds core --iters 5; done
) > gpurun_out/kbench_dcore_nt.log 2>&1
grep -v "^$" gpurun_out/kbench_dcore_nt.log | cut -c1-200
