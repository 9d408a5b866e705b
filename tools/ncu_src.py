"""Top stall-sample SASS lines of an ncu report: python tools/ncu_src.py report.ncu-rep [n]  (run where ncu is installed)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]; hdr = rows[k + 1]; ci = {h: i for i, h in enumerate(hdr)}
        body = []
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            body.append(rows[k]); k += 1
        tot = sum(int(r[ci["# Samples"]]) for r in body if r[ci["# Samples"]].isdigit())
        print("==", name[:80], "samples", tot)
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ci[h]]) for r in body if r[ci[h]].isdigit()) for h in stall_cols}
        print("  by reason:", ", ".join(f"{h[6:]} {100*v/max(tot,1):.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))
        for i in order[:top]:
            r = body[i]; n = int(r[ci["# Samples"]] or 0)
            why = max(stall_cols, key=lambda h: int(r[ci[h]] or 0))
            print(f"  {100*n/max(tot,1):5.1f}%  line {i:4d}  {why[6:]:14s} exec {r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:90]}")
    else:
        k += 1
