"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep (needs --import-source on and -lineinfo).
usage: python scripts/ncu_inst_by_line.py report.ncu-rep kernel_substring [top_n] [samples]   (sorted by instructions, or by stall samples)"""
import csv, io, subprocess, sys

rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn, path, hdr, rows, blocks = None, None, None, [], []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": path = r[1]; continue
    if r[0] == "Function Name":
        fn = r[1]; hdr = None; rows = []; blocks.append((fn, path, rows)); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit(): rows.append(r)
iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), 1
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
agg = {}
for fn, path, rows in blocks:
    if want not in fn: continue
    for r in rows:
        k = (path.split("/")[-1], int(r[0]))
        a = agg.setdefault(k, [0, 0, r[iSrc], {}])
        a[0] += int(r[iI]) if r[iI].isdigit() else 0
        a[1] += int(r[iS]) if r[iS].isdigit() else 0
        for i in stall:
            if r[i].isdigit() and int(r[i]): a[3][hdr[i][6:]] = a[3].get(hdr[i][6:], 0) + int(r[i])
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions:", tot, " samples:", tots)
bysmpl = len(sys.argv) > 4 and sys.argv[4] == "samples"
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if bysmpl else 0])[:top]:
    st = sorted(a[3].items(), key=lambda x: -x[1])[:2]
    print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[1] / tots:5.1f}% smpl  {k[0]}:{k[1]:<4d} {a[2].strip()[:90]}  {st}")
