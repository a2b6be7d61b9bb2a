#!/bin/bash
# Round-2 evidence run: bench line, ncu launch list of the bench command, ncu --set full of the main kernels (cfg2 L2) and of
# the generic GEMM (CIFAR (2,23->24)), config-3 grid, config-5 bench lines
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench_cfg2.json 2> gpurun_out/r02f_bench_cfg2.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/kbench.py --layers L2 --kinds fwd,core,input --train --once > gpurun_out/kbench_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_fast_kernel|tc_dcore16_kernel" -c 3 -o gpurun_out/prof_r02f_main python tools/kbench.py --layers L2 --kinds fwd,core,input --train --once > gpurun_out/ncu_main.log 2>&1
echo "ncu main rc=$?"
ncu -i gpurun_out/prof_r02f_main.ncu-rep --page raw --csv > gpurun_out/prof_r02f_main_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02f_main_raw.csv > gpurun_out/r02f_main_kernels_ncu.txt
timeout 300 python tools/kbench.py --layers c23 --batch 64 --kinds fwd,input --train --once > gpurun_out/kbench_once2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -c 2 -o gpurun_out/prof_r02f_generic python tools/kbench.py --layers c23 --batch 64 --kinds fwd,input --train --once > gpurun_out/ncu_generic.log 2>&1
echo "ncu generic rc=$?"
ncu -i gpurun_out/prof_r02f_generic.ncu-rep --page raw --csv > gpurun_out/prof_r02f_generic_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02f_generic_raw.csv > gpurun_out/r02f_generic_kernels_ncu.txt
rm -f gpurun_out/prof_r02f_main.ncu-rep gpurun_out/prof_r02f_generic.ncu-rep
for wl in cfg5_chain cfg5_convsbs; do timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r02f_bench_$wl.json 2> gpurun_out/r02f_bench_$wl.err; echo "$wl rc=$?"; done
timeout 1200 python benchmarks/eps_microbench.py --json gpurun_out/r02f_microbench_cfg3.json > gpurun_out/r02f_microbench_cfg3.txt 2>&1; echo "microbench rc=$?"
tail -5 gpurun_out/r02f_microbench_cfg3.txt
python - <<'P'
import json
for n in ("cfg2","cfg5_chain","cfg5_convsbs","reference"):
    try:
        d=json.load(open(f"gpurun_out/r02f_bench_{n}.json")); print(n, d.get("metric"), d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(n, "failed", e)
P
