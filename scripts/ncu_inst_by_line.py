"""Per-source-line instruction counts of one kernel in an .ncu-rep (needs --import-source on and -lineinfo).
usage: python scripts/ncu_inst_by_line.py report.ncu-rep kernel_substring [top_n]"""
import csv, io, subprocess, sys

rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; kern.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
for k in kern:
    if want not in k["name"]: continue
    h = k["hdr"]
    print("kernel:", k["name"][:80]); print("columns:", h[:12])
    iI = h.index("# Instructions Executed") if "# Instructions Executed" in h else None
    iSrc = h.index("Source")
    if iI is None:
        cand = [c for c in h if "Instructions Executed" in c]; print("cands", cand); iI = h.index(cand[0])
    tot = sum(int(r[iI]) for r in k["rows"] if r[iI].isdigit())
    print("total warp instructions:", tot)
    for r in sorted(k["rows"], key=lambda r: -int(r[iI]) if r[iI].isdigit() else 0)[:top]:
        print(f"{100 * int(r[iI]) / tot:5.1f}%  {int(r[iI]):>10d}  {r[iSrc].strip()[:110]}")
    break
