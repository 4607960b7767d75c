"""Markdown summary of an `ncu --set full` report: one table per kernel (first launch of each) + top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNNx_summary.md"""
import csv, subprocess, sys, io
rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu capture"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
print(f"# {title}\n")
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    if name in seen:
        continue
    seen.add(name)
    print(f"## {name}\n\n| metric | value |\n|---|---|")
    for w in want:
        if w in idx:
            print(f"| {w} | {r[idx[w]]} {units[idx[w]]} |")
    st = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("| top stall reasons (share of stall cycles per issue) | " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:6]) + " |\n")
