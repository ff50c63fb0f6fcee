#!/usr/bin/env python
"""profiles/ncu_traffic_r1.json (bench.py's roofline.traffic) from a --set full capture: DRAM bytes read / written
and duration per launch, keyed by kernel name.  usage: ncu_traffic.py X.ncu-rep out.json "<shape note>" """
import csv, io, json, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v * scale
kern = {}
for r in data:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    kern.setdefault(name, []).append({"dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                                      "time_us": val(r, "gpu__time_duration.sum")})
json.dump({"shape": note, "source": rep.split("/")[-1] + " (ncu --set full, one step)", "kernels": kern}, open(out, "w"), indent=1)
print("wrote", out, {k: len(v) for k, v in kern.items()})
