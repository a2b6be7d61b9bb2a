#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( for st in 0 1; do echo "== STAGGER=$st"; DCTN_B200_STAGGER=$st timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd --train --iters 7; done
  echo "== again STAGGER=0"; DCTN_B200_STAGGER=0 timeout 300 python tools/kbench.py --layers L2 --kinds fwd --train --iters 7
  echo "== again STAGGER=1"; DCTN_B200_STAGGER=1 timeout 300 python tools/kbench.py --layers L2 --kinds fwd --train --iters 7
) > gpurun_out/kbench_stagger.log 2>&1
grep -v "^$" gpurun_out/kbench_stagger.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "tc or model or full_size" > gpurun_out/pytest_tc.log 2>&1; tail -2 gpurun_out/pytest_tc.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_g.json 2> gpurun_out/bench_cfg2_g.err; tail -2 gpurun_out/bench_cfg2_g.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_cfg2_g.json"))
print("cfg2", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]))
for k in d["roofline"]["all_kernels"]: print("    ",k)
P
timeout 300 python tools/kbench.py --layers L2 --kinds fwd,core,input --train --once > gpurun_out/kbench_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_fast_kernel|tc_dcore16_kernel" -c 3 -o gpurun_out/prof_r02_main python tools/kbench.py --layers L2 --kinds fwd,core,input --train --once > gpurun_out/ncu_main.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_main.log
