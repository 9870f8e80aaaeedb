"""Scratch: per-pass timing of the 3-D r2c (z, y, x) and variants, CUDA events"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodykit_b200 import _lib
from nbodykit_b200._lib import lib, check
from nbodykit_b200.pmesh.pm import ParticleMesh, RealField, ComplexField, _ptr, _stream, _CODE
from nbodykit_b200.comm import SelfComm
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dt = sys.argv[2] if len(sys.argv) > 2 else "f8"
pm = ParticleMesh(BoxSize=1.0, Nmesh=N, dtype=dt, comm=SelfComm())
r = RealField(pm); r.value.normal_()
c = ComplexField(pm)
code = _CODE[pm.typestr]
Nzc = N // 2 + 1
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
cb = c.value.numel() * c.value.element_size()
z = t(lambda: check(lib().nbk_fft_z_forward(_ptr(r.value), _ptr(c.value), code, N * N, N, _stream())))
y = t(lambda: check(lib().nbk_fft_lines(_ptr(c.value), code, N, Nzc, Nzc, N, N * Nzc, 0, 1.0, _stream())))
x = t(lambda: check(lib().nbk_fft_lines(_ptr(c.value), code, N, N * Nzc, N * Nzc, 1, 0, 0, 1.0, _stream())))
tot = t(lambda: r.r2c(out=c))
print("%d^3 %s [%s]: z %.3f ms (%.0f GB/s)  y %.3f ms (%.0f GB/s)  x %.3f ms (%.0f GB/s)  r2c %.3f ms (4x field / t = %.0f GB/s)" % (
    N, dt, " ".join("%s=%s" % (k[8:], v) for k, v in os.environ.items() if k.startswith("NBK_FFT_")),
    z, (fb + cb) / z / 1e6, y, 2 * cb / y / 1e6, x, 2 * cb / x / 1e6, tot, 4 * fb / tot / 1e6))
# correctness of the line passes against torch.fft on a small slab
rr = RealField(pm); rr.value.copy_(r.value)
got = rr.r2c().value
ref = torch.fft.rfftn(r.value.double()) / N ** 3
err = (got.to(torch.complex128) - ref).abs().max().item() / ref.abs().max().item()
print("   max |r2c - torch.fft.rfftn| / max = %.2e" % err)
