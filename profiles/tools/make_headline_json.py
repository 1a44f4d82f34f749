"""profiles/ncu_headline.json (what bench.py's roofline block reads) from two digests of the headline step:
    python profiles/tools/make_headline_json.py COLD.json WARM.json "source text" > profiles/ncu_headline.json
COLD = default ncu cache control (caches flushed before every pass), WARM = --cache-control none (steady state)."""
import json
import sys

cold, warm = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
out = {"_source": sys.argv[3]}
for key, frags in (("prove_kernel", ("prove_kernel",)), ("verify_kernel", ("verify_log_kernel", "verify_fast_kernel"))):
    frag = next(f for f in frags if any(f in r["kernel"] for r in cold))     # the table-path verifier when the capture has it
    c = next(r for r in cold if frag in r["kernel"])
    w = next(r for r in warm if frag in r["kernel"])
    out[key] = {
        "kernel": c["kernel"], "items": c["items"], "duration_us": c["duration_us"], "duration_us_warm": w["duration_us"], "registers": c["registers"],
        "imad_thread_inst_per_item": c["imad_thread_inst_per_item"], "thread_inst_per_item": c["thread_inst_per_item"],
        "warp_instructions_per_launched_warp": c["warp_instructions"] / (c["grid"] * c["block"] / 32),
        "avg_active_threads_per_instruction": c["avg_active_threads_per_instruction"],
        "issue_slot_utilization": c["issue_slot_utilization"], "fmaheavy_pipe_active": c["fmaheavy_pipe_active"],
        "alu_pipe_utilization": c["alu_pipe_utilization"], "lsu_pipe_utilization": c["lsu_pipe_utilization"],
        "shared_bank_conflicts": c["shared_bank_conflicts"], "isetp_sel_per_launched_warp": c.get("isetp_sel_per_warp"),
        "sts_stall_share": c.get("sts_stall_share"), "stall_cycles_per_issue": c["stall_cycles_per_issue"],
        "dram_bytes_per_launch": c["dram_bytes_per_launch"], "dram_bytes_per_launch_warm": w["dram_bytes_per_launch"],
        "opcode_per_launched_warp": c.get("opcode_per_warp"),
    }
print(json.dumps(out, indent=1))
