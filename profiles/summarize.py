"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under
profiles/ (run from the repo root on the CPU box: python profiles/summarize.py r01)."""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []

def us(row):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)

path = "gpurun_out/launches_%s.csv" % tag
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict(); tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        a = agg.setdefault(name, [0.0, 0]); t = us(r); a[0] += t; a[1] += 1; tot += t
    out.append("## Launch list of ONE training step (B=16, N=8192), `ncu --metrics gpu__time_duration.sum "
               "--clock-control none`\n\nper-launch times are cold-cache and serialised: compare SHARES.\n\n"
               "total %.1f us over %d launches\n\n| kernel | us | launches | share |\n|---|---:|---:|---:|" % (tot, len(rows)))
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append("| `%s` | %.1f | %d | %.1f%% |" % (k[:70], t, n, 100 * t / tot))
    out.append("")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
traffic = {}
for rep in sorted(f for f in os.listdir("gpurun_out") if f.endswith("_%s.ncu-rep" % tag)):
    txt = subprocess.run(["ncu", "-i", os.path.join("gpurun_out", rep), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out.append("## `ncu --set full` capture: %s\n" % rep)
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        out.append("kernel `%s` grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|" % (
            d.get("Kernel Name", "")[:90], d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                out.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        try:
            def b(x, u): return float(x.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            t = b(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")]) + \
                b(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
            traffic.setdefault(rep.split("_")[0], []).append(t)
            out.append("| dram traffic (read+write) | %.0f | byte |" % t)
        except Exception:
            pass
        out.append("")
open("profiles/ncu_summary_%s.md" % tag, "w").write("\n".join(out) + "\n")
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open("profiles/traffic_%s.json" % tag, "w"), indent=1)
print("wrote profiles/ncu_summary_%s.md" % tag)
