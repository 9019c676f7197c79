"""Summarise an .ncu-rep (ncu --set full --import-source on) into JSON + the top stall lines per kernel.
usage: python scripts/ncu_summary.py report.ncu-rep > profiles/<name>.summary.txt"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
units = dict(zip(hdr, rows[1]))
print("# ncu summary of", rep)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("\n## kernel:", d.get("Kernel Name", "")[:100])
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {units.get(k, '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:60], "hdr": None, "rows": []}; kern.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen: continue
    seen.add(k["name"])
    h = k["hdr"]; iS = h.index("# Samples"); iSrc = h.index("Source")
    tot = sum(int(r[iS]) for r in k["rows"] if r[iS].isdigit()) or 1
    cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    agg = {}
    for r in k["rows"]:
        for i in cols:
            if r[i].isdigit(): agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print(f"\n## stall sampling: {k['name']}  ({tot} samples)")
    print("  totals:", ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for v, n in sorted(((v, n) for n, v in agg.items()), reverse=True)[:8]))
    for r in sorted(k["rows"], key=lambda r: -int(r[iS]) if r[iS].isdigit() else 0)[:12]:
        st = sorted([(int(r[i]), h[i][6:]) for i in cols if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:2]
        print(f"  {100 * int(r[iS]) / tot:5.1f}%  {r[iSrc].strip()[:64]:64s} {st}")
