"""Stall samples of one kernel aggregated by CUDA source line (needs -lineinfo and --import-source on).
usage: ncu_lines.py file.ncu-rep <kernel substring> [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, sub = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if cur is not None and r: cur["rows"].append(r)
for s in secs:
    if sub not in s["name"]: continue
    h = s["rows"][0]; d = s["rows"][1:]
    if "# Samples" not in h: continue
    iS = h.index("# Samples"); iSrc = h.index("Source")
    iFile = h.index("File Name") if "File Name" in h else None
    iLine = h.index("Line") if "Line" in h else (h.index("#") if "#" in h else None)
    agg = defaultdict(int); tot = 0
    for r in d:
        try: n = int(r[iS] or 0)
        except ValueError: continue
        key = (r[iFile].split("/")[-1] if iFile is not None else "", r[iLine] if iLine is not None else "", r[iSrc].strip()[:100])
        agg[key] += n; tot += n
    print("==", s["name"][:60], "samples", tot, "columns", h[:8])
    for k2, n in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        print("  %5.1f%%  %s:%s  %s" % (100.0 * n / max(tot, 1), k2[0], k2[1], k2[2]))
    break
