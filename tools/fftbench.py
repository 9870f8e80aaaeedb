"""r2c / c2r timing only (scratch)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodykit_b200.pmesh.pm import ParticleMesh, RealField, ComplexField
from nbodykit_b200.comm import SelfComm
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dt = sys.argv[2] if len(sys.argv) > 2 else "f8"
pm = ParticleMesh(BoxSize=1.0, Nmesh=N, dtype=dt, comm=SelfComm())
r = RealField(pm); r.value.normal_()
c = ComplexField(pm)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
fb = r.value.numel() * r.value.element_size()
ms = t(lambda: r.r2c(out=c))
print("r2c %d^3 %s: %.3f ms (6x field bytes moved / t = %.0f GB/s; 4x = %.0f GB/s)" % (N, dt, ms, 6 * fb / ms / 1e6, 4 * fb / ms / 1e6))
ms = t(lambda: c.c2r(out=r))
print("c2r %d^3 %s: %.3f ms" % (N, dt, ms))
