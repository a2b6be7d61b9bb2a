#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream_k2q2 or pixels or eps_vs_oracle or golden or cfg1 or direct" > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_stream.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_stream.log | cut -c1-300 | tail -12
( for rh in 4 9; do for pf in 0 1; do for cap in 0 8; do echo "== RH=$rh PFT=$pf CAP=$cap"; DCTN_B200_STREAM_CAP=$cap DCTN_B200_STREAM_RH=$rh DCTN_B200_STREAM_PFT=$pf timeout 300 python tools/stream_probe.py 2>&1 | grep "O=[24]: core"; done; done; done ) > gpurun_out/stream_probe.log 2>&1
cut -c1-250 gpurun_out/stream_probe.log
