#!/bin/bash
# Final verification of the round: full GPU test-suite, smoke(), default bench + reference arm, launch list, ncu of the CIFAR (2,12->24) kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/pytest_final.log | cut -c1-300 | tail -6
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/r02h_bench_cfg2.json 2> gpurun_out/r02h_bench_cfg2.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/kbench.py --layers c12 --batch 64 --kinds fwd,input --train --once > gpurun_out/kbench_once3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|loo1_rows_kernel" -c 3 -o gpurun_out/prof_r02h_c12 python tools/kbench.py --layers c12 --batch 64 --kinds fwd,input --train --once > gpurun_out/ncu_c12.log 2>&1
echo "ncu c12 rc=$?"
ncu -i gpurun_out/prof_r02h_c12.ncu-rep --page raw --csv > gpurun_out/prof_r02h_c12_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02h_c12_raw.csv > gpurun_out/r02h_c12_kernels_ncu.txt
rm -f gpurun_out/prof_r02h_c12.ncu-rep
python - <<'P'
import json
for n in ("cfg2","reference"):
    try:
        d=json.load(open(f"gpurun_out/r02h_bench_{n}.json")); print(n, d.get("metric"), round(d.get("value"),1), d.get("unit"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e: print(n, "failed", e)
P
grep -E "kernel:|gpu__time|tensor_cycles|issue_active.avg|dram__bytes" gpurun_out/r02h_c12_kernels_ncu.txt
