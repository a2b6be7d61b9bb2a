#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"loo1_warp_kernel|loo2_from_saved" -c 2 -o gpurun_out/prof_loo python tools/kbench.py --layers k3q3 --batch 512 --kinds input --train --once > gpurun_out/ncu_loo.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_loo.ncu-rep --page raw --csv > gpurun_out/prof_loo_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_loo_raw.csv
python - <<'P'
import csv
rows=list(csv.reader(open("gpurun_out/prof_loo_raw.csv")))
hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d.get("Kernel Name","?")[:60])
    for k,v in d.items():
        if ("stalled" in k and "per_issue_active" in k) or k in ("smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__cycles_active.avg","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__waves_per_multiprocessor","smsp__thread_inst_executed_per_inst_executed.ratio"):
            try:
                if float(v.replace(",",""))>0.3: print("   ",k,v)
            except: pass
P
rm -f gpurun_out/prof_loo.ncu-rep
