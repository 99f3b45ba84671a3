"""Summarise ncu reports (read here, no GPU): `python tools/ncu_summary.py <title> <report.ncu-rep> [...] > profiles/x.md`.
Also prints a JSON line {kernel: dram bytes per launch} on stderr for profiles/ncu_traffic.json."""
import csv, io, json, subprocess, sys
WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
title = sys.argv[1]
print(f"# {title}\n")
traffic = {}
for rep in sys.argv[2:]:
    rows = rows_of(rep)
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    print(f"Report `{rep.split('/')[-1]}`\n")
    print("| metric | unit | " + " | ".join(n[:60] for n in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    def tobytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for n, r in zip(names, data):
        key = n.split("(")[0].replace("void ", "").split("<")[0].replace("reni::", "")
        traffic.setdefault(key, tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]))
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for n, r in zip(names, data):
        top = sorted(((float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h, i in stall), reverse=True)[:5]
        print(f"\nTop stall reasons (warps per issue) {n[:60]}: " + ", ".join(f"{h} {v:.2f}" for v, h in top))
    print()
print(json.dumps({"bytes_per_launch": traffic}), file=sys.stderr)
