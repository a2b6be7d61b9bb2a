#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream_k2q2 or pixels or eps_vs_oracle or golden or cfg1" > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_stream.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_stream.log | cut -c1-300 | tail -12
( for pf in 1024; do echo "== PF=$pf"; DCTN_B200_STREAM_PF=$pf timeout 300 python tools/stream_probe.py; done ) > gpurun_out/stream_probe.log 2>&1
cut -c1-250 gpurun_out/stream_probe.log
