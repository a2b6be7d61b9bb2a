#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lopsided or cifar or tch3 or tc_forward or saved" > gpurun_out/pytest_lop.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lop.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_lop.log | cut -c1-300 | tail -12
( for m in 2 3; do echo "== SPLIT_M=$m"; DCTN_B200_SPLIT_M=$m timeout 300 python tools/kbench.py --layers c6,c12,c23 --batch 64 --kinds fwd,core,input --iters 5; echo "-- train"; DCTN_B200_SPLIT_M=$m timeout 300 python tools/kbench.py --layers c6,c12,c23 --batch 64 --kinds fwd,input --train --iters 5; done ) > gpurun_out/kbench_split.log 2>&1
grep -v "^$" gpurun_out/kbench_split.log | cut -c1-200
