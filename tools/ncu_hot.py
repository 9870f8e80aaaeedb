"""top stall-sample lines of one kernel from an .ncu-rep (source page).  usage: ncu_hot.py rep kernel_regex [cuda|sass] [top]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
mode = sys.argv[3] if len(sys.argv) > 3 else "cuda"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", mode, "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r != h]
S = h.index("# Samples"); I = h.index("Instructions Executed"); SRC = h.index("Source")
stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(num(r[S]) for r in body); toti = sum(num(r[I]) for r in body)
print("kernel:", rows[0][1][:90] if rows[0] else "", " total samples %d, warp instr %d" % (tot, toti))
xtra = [h.index(k) for k in ("L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global Excessive") if k in h]
for r in sorted(body, key=lambda r: -num(r[S]))[:top]:
    st = sorted(((num(r[i]), h[i][6:]) for i in stalls), reverse=True)[:3]
    print("%5.1f%% inst %5.1f%% | %-22s | %s | %s" % (100 * num(r[S]) / max(tot, 1), 100 * num(r[I]) / max(toti, 1),
          " ".join("%s:%d" % (n, v) for v, n in st if v > 0)[:22], ",".join(r[i] for i in xtra), r[SRC].strip()[:110]))
