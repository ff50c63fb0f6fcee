"""Stall samples of k_downdate_ws2 split by warp role (producer / consumer / epilogue / barrier spin loops).
Role boundaries are found from the SASS itself: EXIT instructions separate the roles, the out-of-line
mbarrier spin loops sit after the last EXIT in the order empty, full, pfull, cfull.
usage: ncu_roles.py file.ncu-rep"""
import csv, io, subprocess, sys
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if cur is not None and r: cur["rows"].append(r)
seen = set()
for s in secs:
    if "k_downdate_ws2" not in s["name"]: continue
    h = s["rows"][0]; d = s["rows"][1:]
    iS, iSrc = h.index("# Samples"), h.index("Source")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[iS] or 0) for r in d)
    if (s["name"], tot) in seen: continue
    seen.add((s["name"], tot))
    exits = [i for i, r in enumerate(d) if "EXIT" in r[iSrc]]
    waits = [i for i, r in enumerate(d) if "TRYWAIT" in r[iSrc]]
    # layout: setup | producer (ends at 2nd/3rd EXIT) | consumer | epilogue | spin loops
    last_exit = exits[-1]
    spins = [w for w in waits if w > last_exit]
    # role starts: after EXIT #1 (early exit of setup) producer; consumer starts after the EXIT that follows the producer's
    # UBLKCPs; epilogue after the consumer's EXIT
    ub = [i for i, r in enumerate(d) if "UBLKCP" in r[iSrc]]
    prod_end = min(e for e in exits if e > ub[1])          # producer has the first two UBLKCPs
    cons_end = min(e for e in exits if e > [w for w in waits if w > prod_end][1])   # consumer: full + pfull waits
    bounds = [("setup", 0, exits[0] + 1), ("producer", exits[0] + 1, prod_end + 2), ("consumer", prod_end + 2, cons_end + 1),
              ("epilogue", cons_end + 1, last_exit + 1)]
    names = ["spin empty (producer idle)", "spin full (consumers wait for W)", "spin pfull (consumers wait for P)", "spin cfull (epilogue idle)"]
    for n, a in zip(names, spins):
        nxt = min([x for x in spins if x > a] + [len(d)])
        bounds.append((n, a, nxt))
    print("==", s["name"][:44], "samples", tot)
    for n, a, b2 in bounds:
        sel = d[a:b2]
        ss = sum(int(r[iS] or 0) for r in sel)
        st = {h[i]: sum(int(r[i] or 0) for r in sel) for i in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        print("  %-36s %5.1f%%  " % (n, 100.0 * ss / tot), ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in top if v))
