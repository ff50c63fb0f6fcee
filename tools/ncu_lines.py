"""Stall samples / executed instructions of one kernel aggregated by CUDA source line (needs -lineinfo and --import-source on).
usage: ncu_lines.py file.ncu-rep [top_n]   (the report should hold one kernel)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; lines = []; fname = ""
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if r and r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    lines.append([fname + ":" + r[0]] + r[1:])
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
names = hdr
tot_s = sum(num(r[iS]) for r in lines); tot_i = sum(num(r[iI]) for r in lines)
print("samples", tot_s, "warp instructions", tot_i)
print("-- by samples")
for r in sorted(lines, key=lambda r: -num(r[iS]))[:topn]:
    # the per-reason columns follow the fixed ones; show the top two reasons
    reasons = sorted(((num(r[i]), names[i]) for i in range(32, min(len(r), len(names)))), reverse=True)[:2]
    print("  %5.1f%% smp %5.1f%% inst  %-18s %-90s %s" % (100.0 * num(r[iS]) / max(tot_s, 1), 100.0 * num(r[iI]) / max(tot_i, 1), r[0], r[1].strip()[:90],
          ", ".join("%s %d" % (n.replace("stall_", ""), c) for c, n in reasons if c)))
