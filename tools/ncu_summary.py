"""Condense one captured kernel of an .ncu-rep into the JSON summary kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep out.json ["command line that produced it"]"""
import csv, io, json, subprocess, sys
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
out = {}
for i, n in enumerate(h):
    if n in KEEP or ("issue_stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n and float(v[i] or 0) >= 0.05):
        out[n] = {"value": v[i], "unit": u[i]}
out["_kernel"] = v[h.index("Kernel Name")]
if len(sys.argv) > 3:
    out["_command"] = sys.argv[3]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1)[:400])
