"""per-launch summary of an ncu --csv launch list: time, DRAM read/write.  usage: ncu_summary.py file.csv [max_rows]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ii = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size")}
d = {}
for r in rows[1:]:
    k = (int(r[ii["ID"]]), r[ii["Kernel Name"]][:48], r[ii["Grid Size"]], r[ii["Block Size"]])
    d.setdefault(k, {})[r[ii["Metric Name"]]] = float(r[ii["Metric Value"]].replace(",", ""))
tot = [0, 0, 0]
for k in sorted(d)[:int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9]:
    m = d[k]
    t, rd, wr = m.get("gpu__time_duration.sum", 0) / 1e6, m.get("dram__bytes_read.sum", 0) / 1e9, m.get("dram__bytes_write.sum", 0) / 1e9
    tot[0] += t; tot[1] += rd; tot[2] += wr
    print("%3d %-48s %-14s %-12s t=%.3f ms R=%.3f GB W=%.3f GB" % (k[0], k[1], k[2], k[3], t, rd, wr))
print("total t=%.3f ms R=%.3f GB W=%.3f GB" % tuple(tot))
