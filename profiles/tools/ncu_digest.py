"""Digest an .ncu-rep (ncu --set full --import-source on) into a text summary and a JSON record per kernel launch.

    python profiles/tools/ncu_digest.py REPORT.ncu-rep [--items N[,N2,...]] [--json OUT.json] > summary.txt

Raw page: duration, registers, occupancy, issue-slot utilisation, pipe utilisation (ALU / FMA / FMA-heavy / LSU), DRAM bytes,
shared-memory bank conflicts, stall reasons per issued instruction.  Source page: executed warp- and thread-level
instructions per SASS opcode (the IMAD family summed separately: that is the numerator of bench.py's roofline.frac),
stall-sample share per opcode, shared-memory wavefronts in excess of the ideal.  --items gives the number of items each
launch processed (in launch order; the last value repeats), so that per-item figures can be printed.
"""
import argparse
import collections
import csv
import json
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
       "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return None if v is None else v * m.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--items", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    items = [int(x) for x in a.items.split(",") if x]
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and r:
            cur["rows"].append(r)
    blocks = blocks[0::2] if len(blocks) >= 2 * (len(rows) - 2) else blocks   # the export lists every launch twice (SASS view, source view)
    out = []
    for k, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("pb::", "")
        n_items = (items[k] if k < len(items) else items[-1]) if items else None
        rec = {"kernel": name, "items": n_items}
        print(f"== {name}" + (f"   [{n_items} items]" if n_items else ""))
        for w in RAW:
            if w in ix:
                print(f"   {w:72s} {r[ix[w]]:>18s} {units[ix[w]]}")
        g = lambda key: num(r[ix[key]]) if key in ix else None
        dur = g("gpu__time_duration.sum")
        if dur is not None and units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            dur /= 1e3
        elif dur is not None and units[ix["gpu__time_duration.sum"]] in ("ms", "msecond"):
            dur *= 1e3
        rec.update(duration_us=dur, registers=g("launch__registers_per_thread"), grid=g("launch__grid_size"), block=g("launch__block_size"),
                   issue_slot_utilization=(g("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100,
                   fmaheavy_pipe_active=(g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0) / 100,
                   fma_pipe_cycles_active=(g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") or 0) / 100,
                   alu_pipe_utilization=(g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") or 0) / 100,
                   lsu_pipe_utilization=(g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active") or 0) / 100,
                   warp_instructions=g("smsp__inst_executed.sum"), avg_active_threads_per_instruction=g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   shared_bank_conflicts=g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                   dram_bytes_per_launch=(to_bytes(g("dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]]) or 0)
                   + (to_bytes(g("dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]]) or 0),
                   dram_throughput_pct=g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
        st = sorted(((num(r[ix[h]]) or 0, h) for h in stalls), reverse=True)
        rec["stall_cycles_per_issue"] = {}
        for v, h in st:
            if v >= 0.05:
                key = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                rec["stall_cycles_per_issue"][key] = round(v, 3)
                print(f"   stall {key:66s} {v:18.3f} warp-cycles per issue")
        if k < len(blocks) and blocks[k]["hdr"]:
            h = {x: i for i, x in enumerate(blocks[k]["hdr"])}
            ops, thr, samples, excess = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
            def col(row, name):     # a column is absent when the launch produced nothing for it (no samples, no shared accesses)
                v = row[h[name]] if name in h and h[name] < len(row) else ""
                return int(v) if v.isdigit() else 0
            for row in blocks[k]["rows"]:
                if "Instructions Executed" not in h or len(row) <= h["Instructions Executed"] or not row[h["Instructions Executed"]].isdigit():
                    continue
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[h["Source"]])
                base = (m.group(2) if m else row[h["Source"]].strip()).split(".")[0]
                ops[base] += col(row, "Instructions Executed")
                thr[base] += col(row, "Thread Instructions Executed")
                samples[base] += col(row, "# Samples")
                excess[base] += col(row, "L1 Wavefronts Shared Excessive")
            tot, tots = sum(ops.values()), max(sum(samples.values()), 1)
            warps = (rec["grid"] or 0) * (rec["block"] or 0) / 32 or 1
            print(f"   -- executed SASS mix: {tot} warp instructions, {tot / warps:.1f} per launched warp, {len(blocks[k]['rows'])} static instructions")
            for op, v in ops.most_common(14):
                print(f"      {op:10s} {v / warps:9.1f} /warp {100 * v / tot:5.1f}%   thread-level {thr[op]:>14d}   stall samples {100 * samples[op] / tots:5.1f}%"
                      + (f"   excess smem wavefronts {excess[op]}" if excess[op] else ""))
            rec["imad_thread_inst"] = thr["IMAD"]
            rec["thread_inst"] = sum(thr.values())
            rec["isetp_sel_per_warp"] = (ops["ISETP"] + ops["SEL"]) / warps
            rec["sts_stall_share"] = samples["STS"] / tots
            rec["opcode_per_warp"] = {op: round(v / warps, 1) for op, v in ops.most_common(16)}
            if n_items:
                rec["imad_thread_inst_per_item"] = thr["IMAD"] / n_items
                rec["thread_inst_per_item"] = sum(thr.values()) / n_items
                rec["dram_bytes_per_item"] = rec["dram_bytes_per_launch"] / n_items
                print(f"   -- per item: {rec['thread_inst_per_item']:.1f} thread instructions, {rec['imad_thread_inst_per_item']:.1f} IMAD*, "
                      f"{rec['dram_bytes_per_item']:.1f} DRAM bytes")
        out.append(rec)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
