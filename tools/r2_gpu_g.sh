# round 2, final code: cfg4 (N=500, B=8) evidence - launch list of the whole run and ncu --set full of the downdate and the
# blocked-64 Cholesky kernels (tensor-pipe %, DRAM bytes)
mkdir -p gpurun_out
python tools/cfg4_once.py 8 > gpurun_out/r2g_cfg4_plain.log 2>&1 && tail -1 gpurun_out/r2g_cfg4_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_cfg4_launches.csv python tools/cfg4_once.py 8 > gpurun_out/r2g_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_downdate|k_cb_|k_gemm" -s 60 -c 60 -o /tmp/r2g_cfg4 python tools/cfg4_once.py 8 > gpurun_out/r2g_ncu_full.log 2>&1
tail -n 2 gpurun_out/r2g_ncu_full.log
python tools/ncu_summary.py /tmp/r2g_cfg4.ncu-rep gpurun_out/ncu_r2g_cfg4_summary.md > gpurun_out/r2g_summary.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2g_cfg4_launches.csv", errors="ignore")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    v = float(r[14].replace(",", "")); u = r[13]
    us = v / 1e3 if u in ("nsecond", "ns") else v if u in ("usecond", "us") else v * 1e3
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/launches_r2g_cfg4_summary.md", "w") as f:
    f.write("# cfg4 (N=500, n=3013, B=8), final round-2 code: every launch of tools/cfg4_once.py (2 warm + 2 timed steps), ncu --metrics gpu__time_duration.sum --clock-control none\n")
    f.write("# per-launch times are cold-cache and SERIALISED: compare shares, not absolutes\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
print(open("gpurun_out/launches_r2g_cfg4_summary.md").read()[:1500])
PY
