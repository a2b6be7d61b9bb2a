#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/diag_layers.py "((4,4),(3,12),(2,24))" 2 28 9 102 > gpurun_out/diag_three.log 2>&1; tail -20 gpurun_out/diag_three.log
DCTN_B200_NO_FAST=1 timeout 600 python tools/diag_layers.py "((4,4),(3,12),(2,24))" 2 28 9 102 > gpurun_out/diag_three_nofast.log 2>&1; tail -20 gpurun_out/diag_three_nofast.log
timeout 600 python tools/diag_layers.py "((2,23),(2,24))" 4 32 5 126 > gpurun_out/diag_cifar23.log 2>&1; tail -12 gpurun_out/diag_cifar23.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tc_backward_core or full_size or range_norm or tch3" > gpurun_out/pytest_dcore.log 2>&1; tail -5 gpurun_out/pytest_dcore.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_dcorefold.json 2> gpurun_out/bench_cfg2_dcorefold.err; tail -2 gpurun_out/bench_cfg2_dcorefold.err
