#!/usr/bin/env python
"""Print the key counters of an .ncu-rep (first kernel): python profiles/ncu_keys.py <file.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
pats = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_uniform",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum ",
        "launch__shared_mem_per_block", "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__ctas_launched"]
for i, h in enumerate(hdr):
    if any(h.startswith(p.strip()) for p in pats):
        try:
            v = float(vals[i].replace(",", ""))
            if "stalled" in h and v < 0.2:
                continue
        except ValueError:
            pass
        print(f"{h:92s} {vals[i]:>18s} {units[i]}")
