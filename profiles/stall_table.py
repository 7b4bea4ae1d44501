"""ncu --page source --csv of a .ncu-rep -> the SASS instructions with the most warp-stall samples, grouped by
the role region of the kernel they fall in (profiles/README_*.md cites the output).
    python profiles/stall_table.py gpurun_out/tcgemm_r02.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
# first launch only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["# Samples"]) for r in rows)
print("kernel:", lines[start - 1][:120])
print("total samples %d over %d SASS instructions" % (tot, len(rows)))
cols = [c for c in rows[0].keys() if c.startswith("stall_") or "Stall" in c]
rs = sorted(rows, key=lambda r: -int(r["# Samples"]))[:top]
print("| samples | share | SASS | dominant stall reasons |")
print("|---:|---:|---|---|")
reason_cols = [c for c in rows[0].keys() if c not in ("Address", "Source") and not c.startswith(("L1", "L2", "Warp Stall", "#", "Instructions", "Thread", "Predicated", "Avg", "Divergent", "Address Space", "Access"))]
for r in rs:
    n = int(r["# Samples"])
    why = sorted(((int(r[c]), c) for c in reason_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print("| %d | %.1f%% | `%s` | %s |" % (n, 100.0 * n / tot, r["Source"].strip()[:70], ", ".join("%s %d" % (c, v) for v, c in why)))
