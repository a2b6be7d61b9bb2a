#!/usr/bin/env python
"""Condenses an `ncu --page raw --csv` dump of ONE kernel launch into the handful of metrics DESIGN.md / bench.py cite.
   ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/xxx.txt"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"# kernel: {d.get('Kernel Name', '?')}   grid {d.get('launch__grid_size', '?')}  block {d.get('launch__block_size', '?')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"{k:110s} {d[k]:>18s} {u.get(k, '')}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
