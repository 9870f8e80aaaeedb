"""nvcc -Xptxas -v summary: demangled kernel name, registers, spills, smem.  usage: ptxas_report.py file.cu [regex] [-- extra nvcc flags]"""
import re, subprocess, sys
src = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "--" else "."
extra = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else []
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v", "-c", src, "-o", "/tmp/_ptxas_report.o"] + extra
out = subprocess.run(cmd, capture_output=True, text=True).stderr
names = re.findall(r"Compiling entry function '(\S+)'", out)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
blocks = out.split("Compiling entry function")[1:]
for n, d, b in zip(names, dem, blocks):
    if not re.search(pat, d):
        continue
    regs = re.search(r"Used (\d+) registers", b)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    smem = re.search(r"(\d+) bytes smem", b)
    d = re.sub(r"\(.*", "", d)
    print("%-70s regs=%s spill=%s/%s smem=%s" % (d[-70:], regs.group(1) if regs else "?", spill.group(1) if spill else "?", spill.group(2) if spill else "?", smem.group(1) if smem else "0"))
if "error" in out:
    print(out[-3000:])
