"""Stall samples / executed instructions of each kernel in a report, aggregated by CUDA source line
(needs -lineinfo and --import-source on).  usage: ncu_lines.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))


def num(x):
    try: return int(x)
    except ValueError: return 0


kernels = []   # [name, hdr, lines]
cur = None; fname = ""
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        if cur is None or cur[0] != r[1] or cur[3]: cur = [r[1], None, [], False]; kernels.append(cur)
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        if cur is not None and cur[2]: cur[3] = False
        continue
    if r[0] == "Line No":
        if cur is not None: cur[1] = r
        continue
    if cur is None or cur[1] is None or not r[0].isdigit(): continue
    cur[2].append([fname + ":" + r[0]] + r[1:])
# sections of one kernel launch repeat the function name per file: merge consecutive sections with the same name
merged = []
for k in kernels:
    if merged and merged[-1][0] == k[0] and k[1] == merged[-1][1]: merged[-1][2].extend(k[2])
    else: merged.append(k)
for name, hdr, lines, _ in merged:
    if not hdr or not lines: continue
    iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
    tot_s = sum(num(r[iS]) for r in lines); tot_i = sum(num(r[iI]) for r in lines)
    print("==", name[:100]); print("samples", tot_s, "warp instructions", tot_i)
    for r in sorted(lines, key=lambda r: -num(r[iS]))[:topn]:
        reasons = sorted(((num(r[i]), hdr[i]) for i in range(32, min(len(r), len(hdr)))), reverse=True)[:2]
        print("  %5.1f%% smp %5.1f%% inst  %-18s %-90s %s" % (100.0 * num(r[iS]) / max(tot_s, 1), 100.0 * num(r[iI]) / max(tot_i, 1), r[0], r[1].strip()[:90],
              ", ".join("%s %d" % (n.replace("stall_", ""), c) for c, n in reasons if c)))
