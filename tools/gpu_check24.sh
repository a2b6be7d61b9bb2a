#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "leave_one_out_rows or generic or cifar or three_layer or core or dcore or cfg2" > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reg.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_reg.log | cut -c1-300 | tail -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3q3.csv python tools/kbench.py --layers k3q3 --batch 512 --kinds core --once > gpurun_out/ncu_l_k3q3.log 2>&1
grep -E "build_tables16|tc_dcore16" gpurun_out/launches_k3q3.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
