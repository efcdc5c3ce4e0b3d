"""Turn one round's ncu outputs into the tracked summaries under profiles/.

    python scripts/ncu_summary.py <tag>        e.g. r01h  (reads gpurun_out/<tag>_prof.ncu-rep,
                                                gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_bench.json)

Writes profiles/<tag>_summary.md (launch list shares + the --set full metrics of the dominant kernel)
and profiles/traffic.json (dram bytes per launch, read by bench.py for roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        res.append({k: (d.get(k, ""), units[hdr.index(k)]) for k in KEYS if k in hdr} | {"name": d.get("Kernel Name", "?")})
    return res


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    return [(r[4], r[7], r[8], float(r[14])) for r in rows]      # name, block, grid, ns


def main():
    tag = sys.argv[1]
    g = lambda s: os.path.join(ROOT, "gpurun_out", f"{tag}_{s}")
    lines = [f"# ncu summary {tag}", ""]
    if os.path.exists(g("bench.json")):
        try:
            b = json.loads(open(g("bench.json")).read().strip().splitlines()[-1])
            lines += ["## bench line of the same build (CUDA events, no profiler)", "",
                      f"- value {b['value']:.4g} {b['unit']}, {b['ms_per_step']:.4f} ms/step, roofline frac "
                      f"{b['roofline']['frac']:.4f} of {b['roofline']['peak']} GB/s ({b['roofline']['peak_source']})",
                      f"- e2e {b['e2e']['value']:.4g} {b['unit']}; cpu_baseline {b['cpu_baseline']['value']:.4g} "
                      f"({b['cpu_baseline']['cores']} cores); clocks {b['clocks']}", ""]
        except Exception as ex:  # noqa: BLE001
            lines += [f"(bench line unreadable: {ex})", ""]
    if os.path.exists(g("launches.csv")):
        ll = launch_list(g("launches.csv"))
        tot = sum(x[3] for x in ll) or 1.0
        lines += ["## launch list of the timed region (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
                  "cold-cache, serialised: compare shares, not absolutes", "",
                  "| # | kernel | grid | block | us | share |", "|---|---|---|---|---|---|"]
        for i, (name, blk, grid, ns) in enumerate(ll):
            short = name.split("(")[0].split("::")[-1][:60]
            lines.append(f"| {i} | {short} | {grid} | {blk} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% |")
        lines.append("")
    if os.path.exists(g("prof.ncu-rep")):
        ms = raw_metrics(g("prof.ncu-rep"))
        for m in ms:
            lines += [f"## `ncu --set full` of {m['name'].split('(')[0]}", "", "| metric | value | unit |", "|---|---|---|"]
            for k in KEYS:
                if k in m:
                    lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
            lines.append("")
        m = ms[0]
        conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        try:
            tr = sum(float(m[k][0].replace(",", "")) * conv[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            json.dump({"dram_bytes_per_launch": tr, "source": f"profiles/{tag}_summary.md (ncu --set full, one launch)",
                       "kernel": m["name"].split("(")[0]}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
            lines += [f"dram traffic per launch: {tr / 1e9:.3f} GB", ""]
        except Exception as ex:  # noqa: BLE001
            lines += [f"(traffic not derived: {ex})", ""]
    open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
