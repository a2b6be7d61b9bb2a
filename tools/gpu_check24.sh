#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "generic or cifar or three_layer or tc_forward_and_input or tch3 or saved" > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reg.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_reg.log | cut -c1-300 | tail -8
bash tools/gpu_check22.sh 2>&1 | grep -E "^==|loo|sum_slices|gather|tc_gemm_kernel<0|build_tables"
