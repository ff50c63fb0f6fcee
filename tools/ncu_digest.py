"""Digest of an .ncu-rep: per launch headline numbers, stall reasons per issue, opcode mix and the
SASS lines holding most stall samples.  usage: ncu_digest.py file.ncu-rep [top_n]"""
import csv, subprocess, sys, io
from collections import defaultdict
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name): return hdr.index(name) if name in hdr else None
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct"]
for r in data:
    print("==", r[col("Kernel Name")][:40], r[col("Block Size")], r[col("Grid Size")])
    for w in want:
        i = col(w)
        if i is not None: print("   %-70s %s %s" % (w, r[i], units[i]))
    st = [(float(r[i]), hdr[i].split("issue_stalled_")[1].split("_per")[0]) for i in range(len(hdr))
          if "issue_stalled" in hdr[i] and hdr[i].endswith("per_issue_active.ratio") and r[i]]
    print("   stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if cur is not None and r: cur["rows"].append(r)
seen = set()
for s in secs:
    h = s["rows"][0]; d = s["rows"][1:]
    iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    tot = sum(int(r[iS] or 0) for r in d); totE = sum(int(r[iE] or 0) for r in d)
    key = (s["name"], tot)
    if key in seen: continue
    seen.add(key)
    print("== source:", s["name"][:40], "samples", tot, "inst", totE)
    hE = defaultdict(int); hS = defaultdict(int)
    for r in d:
        t = r[iSrc].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        hE[op] += int(r[iE] or 0); hS[op] += int(r[iS] or 0)
    print("   opcodes:", ", ".join("%s %.1f%%/%.1f%%" % (op, 100 * v / totE, 100 * hS[op] / tot) for op, v in sorted(hE.items(), key=lambda kv: -kv[1])[:14]), "(exec/samples)")
    for r in sorted(d, key=lambda r: -int(r[iS] or 0))[:topn]:
        print("   %5.1f%%  %s  %s" % (100 * int(r[iS] or 0) / tot, r[h.index("Address")][-5:], r[iSrc].strip()[:90]))
