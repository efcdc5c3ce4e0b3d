"""CPU: one `ncu --set full` report -> a tracked markdown summary (key metrics + warp stall shares).
    python scripts/ncu_kernel_summary.py gpurun_out/<tag>_bwd.ncu-rep profiles/<tag>_bwd_summary.md "what was run"
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
rep, out, what = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
lines = [f"# ncu --set full: {d.get('Kernel Name', ('?', ''))[0].split('(')[0]}", "", f"`{rep}` -- {what}", "", "| metric | value | unit |", "|---|---|---|"]
for k in KEYS:
    if k in d and d[k][0]:
        lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
st = {}
for k in d:
    if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and d[k][0]:
        try:
            st[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(d[k][0].replace(",", ""))
        except ValueError:
            pass
tot = sum(st.values()) or 1.0
lines += ["", "warp stall reasons (pc sampling, all samples):", "", "| reason | share |", "|---|---|"]
for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
    lines.append(f"| {k} | {100 * v / tot:.1f} % |")
try:
    tr = sum(float(d[k][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[d[k][1]]
             for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    lines += ["", f"dram traffic of this launch: {tr / 1e9:.3f} GB"]
except Exception:
    pass
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
