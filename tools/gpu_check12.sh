#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
( for db in 0 1; do echo "== DBUF=$db"; DCTN_B200_FAST_DBUF=$db timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,input --train --iters 7; done
  echo "== DBUF=0 L1 fwd notrain"; DCTN_B200_FAST_DBUF=0 timeout 300 python tools/kbench.py --layers L1 --kinds fwd --iters 7
  echo "== DBUF=1 L1 fwd notrain"; DCTN_B200_FAST_DBUF=1 timeout 300 python tools/kbench.py --layers L1 --kinds fwd --iters 7
  echo "== cifar generic (8 producers)"; timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,core,input --train --iters 3
) > gpurun_out/kbench_dbuf.log 2>&1
grep -v "^$" gpurun_out/kbench_dbuf.log | cut -c1-200
timeout 300 python tools/bias_probe.py > gpurun_out/bias_probe3.log 2>&1; tail -12 gpurun_out/bias_probe3.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_h.json 2> gpurun_out/bench_cfg2_h.err; tail -2 gpurun_out/bench_cfg2_h.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_cfg2_h.json"))
print("cfg2", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]))
for k in d["roofline"]["all_kernels"]: print("    ",k)
P
for wl in cifar_2_12__2_24 cifar_2_23__2_24 three_eps one_eps; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_h.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_h.json')); print('$wl', round(d['value']), 'img/s', round(d['ms_per_step'],3)); [print('    ',k) for k in d['roofline']['all_kernels']]"; done
