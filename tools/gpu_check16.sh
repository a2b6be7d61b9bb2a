#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
timeout 600 python benchmarks/eps_microbench.py --ks 2 --qs 2,3,4 --iters 11 --json gpurun_out/microbench_k2.json > gpurun_out/microbench_k2.log 2>&1; cut -c1-220 gpurun_out/microbench_k2.log
timeout 300 python tools/stream_probe.py > gpurun_out/stream_probe.log 2>&1; cut -c1-250 gpurun_out/stream_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stream_k2q2_d" -c 2 -o gpurun_out/prof_r02_stream_bwd python tools/stream_probe.py --once > gpurun_out/ncu_stream_bwd.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_stream_bwd.log
ncu -i gpurun_out/prof_r02_stream_bwd.ncu-rep --page raw --csv > gpurun_out/prof_r02_stream_bwd_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02_stream_bwd_raw.csv > gpurun_out/prof_r02_stream_bwd_summary.txt
