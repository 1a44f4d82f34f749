"""Summarise an .ncu-rep (raw page) per kernel launch: duration, registers, occupancy, issue utilisation, instruction
counts, DRAM traffic.  usage: python profiles/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    print(f"== {name}")
    for w in want:
        if w in ix:
            print(f"   {w:70s} {r[ix[w]]:>18s} {units[ix[w]]}")
    # warp-cycles spent per issued instruction, by stall reason (largest first)
    st = sorted(((float(r[ix[h]].replace(",", "") or 0), h) for h in stalls), reverse=True)
    for v, h in st:
        if v >= 0.05:
            print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:62s} {v:18.3f} warp-cycles per issue")
