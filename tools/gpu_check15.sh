#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stream_k2q2" -c 4 -o gpurun_out/prof_r02_stream python tools/stream_probe.py --once > gpurun_out/ncu_stream.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_stream.log
ncu -i gpurun_out/prof_r02_stream.ncu-rep --page raw --csv > gpurun_out/prof_r02_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02_stream_raw.csv > gpurun_out/prof_r02_stream_summary.txt
ls -la gpurun_out/
