#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
R1=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_r1.so
( echo "== r1"; DCTN_B200_LIB=$R1 timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 5
  echo "== new"; timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 5
  echo "== r1 cifar"; DCTN_B200_LIB=$R1 timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 3
  echo "== new cifar"; timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 3
  echo "== r1 L1 notrain"; DCTN_B200_LIB=$R1 timeout 300 python tools/kbench.py --layers L1 --kinds fwd --iters 5
  echo "== new L1 notrain"; timeout 300 python tools/kbench.py --layers L1 --kinds fwd --iters 5
) > gpurun_out/kbench_ab2.log 2>&1
grep -v "^$" gpurun_out/kbench_ab2.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_e.json 2> gpurun_out/bench_cfg2_e.err; tail -2 gpurun_out/bench_cfg2_e.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_cfg2_e.json"))
print("cfg2", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]))
for k in d["roofline"]["all_kernels"]: print("    ",k)
P
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 2 ) > gpurun_out/kbench_timing4.log 2>&1
grep -v "^$" gpurun_out/kbench_timing4.log | awk '!seen[$0]++' | cut -c1-420 | tail -30
