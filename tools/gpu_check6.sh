#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest.log | cut -c1-300 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_d.json 2> gpurun_out/bench_cfg2_d.err; tail -2 gpurun_out/bench_cfg2_d.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_cfg2_d.json"))
print("cfg2", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]))
for k in d["roofline"]["all_kernels"]: print("    ",k)
P
for wl in cifar_2_12__2_24 cifar_2_23__2_24 three_eps; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_d.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_d.json')); print('$wl', round(d['value']), 'img/s', round(d['ms_per_step'],3)); [print('    ',k) for k in d['roofline']['all_kernels']]"; done
export DCTN_B200_LIB=$GRAFT_REPO_ROOT/dctn_b200/libdctn_b200_timing.so
export DCTN_TCG_DEBUG=1
( for nt in 64 96 128; do echo "== dcore NT=$nt"; DCTN_B200_DCORE_NT=$nt timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2; done
  echo "== dcore skip-gen NT=96"; DCTN_B200_SKIP_GEN=1 timeout 300 python tools/kbench.py --layers L2 --kinds core --iters 2
  echo "== full"; timeout 300 python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --train --iters 3
) > gpurun_out/kbench_timing3.log 2>&1
grep -v "^$" gpurun_out/kbench_timing3.log | awk '!seen[$0]++' | cut -c1-420 | tail -30
