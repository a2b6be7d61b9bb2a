#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "leave_one_out_rows or generic or cifar or three_layer or ffma or f64 or float64 or small or input" > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reg.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_reg.log | cut -c1-300 | tail -12
for l in c23 k3q3; do
  B=64; [ $l = k3q3 ] && B=512
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$l.csv python tools/kbench.py --layers $l --batch $B --kinds input --train --once > gpurun_out/ncu_l_$l.log 2>&1
  echo "== $l"; grep -E "loo|sum_slices" gpurun_out/launches_$l.csv | awk -F'","' '{print substr($5,1,50), $NF}'
done
