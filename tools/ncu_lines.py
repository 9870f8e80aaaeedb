"""per-SOURCE-LINE instruction counts and stall samples of one kernel: joins the SASS page of an .ncu-rep with the line
table of the same build (nvdisasm -g on the cubin).  usage: ncu_lines.py rep kernel_regex cubin mangled_substring [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, rx, cubin, mangled = sys.argv[1:5]
import os
SRCDIR = os.environ.get("NBK_SRCDIR", "/root/repo/nbodykit_b200/csrc/")
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
skip = sys.argv[6] if len(sys.argv) > 6 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r != h]
body = body[:len(body) // 2] if len(body) % 2 == 0 and body[0] == body[len(body) // 2] else body
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
I, S, SRC = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass instructions: ncu %d, nvdisasm %d" % (len(body), len(lines)))
n = min(len(body), len(lines))
acc = defaultdict(lambda: [0.0, 0.0])
for r, ln in zip(body[:n], lines[:n]):
    acc[ln][0] += num(r[I]); acc[ln][1] += num(r[S])
ti = sum(v[0] for v in acc.values()); ts = sum(v[1] for v in acc.values())
src = {}
for (f, ln), v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    try:
        if f not in src: src[f] = open((SRCDIR + f)).read().splitlines()
        text = src[f][ln - 1].strip()[:100]
    except Exception:
        text = ""
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100 * v[0] / ti, 100 * v[1] / max(ts, 1), f, ln, text))
