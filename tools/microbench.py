"""Stage timings on one GPU (CUDA events, warm-up, inputs >> L2).  Scratch tool, not the bench contract."""
import argparse
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodykit_b200.pmesh.pm import ParticleMesh, RealField, ComplexField
from nbodykit_b200.comm import SelfComm
from nbodykit_b200.algorithms.fftpower import project_to_basis_device


def timeit(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=512)
    ap.add_argument("--npart", type=float, default=1e8)
    ap.add_argument("--dtype", default="f8")
    args = ap.parse_args()
    N, n = args.nmesh, int(args.npart)
    L = 1024.
    dev = torch.device("cuda", 0)
    pm = ParticleMesh(BoxSize=L, Nmesh=N, dtype=args.dtype, comm=SelfComm())
    g = torch.Generator(device=dev); g.manual_seed(1)
    pos = torch.rand((n, 3), device=dev, dtype=torch.float32, generator=g) * L
    out = {"nmesh": N, "npart": n, "dtype": args.dtype}
    real = RealField(pm); real[...] = 0
    for name, p in [("random", pos)]:
        for res in ["cic", "tsc"]:
            for method in ["direct", "tiled"]:
                t = timeit(lambda: pm.paint(p, resampler=res, hold=True, out=real, method=method))
                out["paint_%s_%s_%s_ms" % (res, name, method)] = t
                print("paint %s %s %s: %.3f ms (median %.3f) -> %.3e particles/s" % (res, name, method, t[0], t[1], n / t[0] * 1e3), flush=True)
    # cell-sorted (morton-ish: sort by cell id)
    cell = (pos / (L / N)).floor().long()
    key = (cell[:, 0] * N + cell[:, 1]) * N + cell[:, 2]
    ps = pos[torch.argsort(key)].contiguous()
    del cell, key
    for res in ["cic", "tsc"]:
        for method in ["direct", "tiled"]:
            t = timeit(lambda: pm.paint(ps, resampler=res, hold=True, out=real, method=method))
            out["paint_%s_cellsorted_%s_ms" % (res, method)] = t
            print("paint %s cellsorted %s: %.3f ms -> %.3e particles/s" % (res, method, t[0], n / t[0] * 1e3), flush=True)
    r2 = RealField(pm); r2[...] = 0
    for method in ["direct", "tiled"]:
        t = timeit(lambda: pm.paint_interlaced(ps, None, "tsc", real, r2, method=method))
        print("paint tsc interlaced cellsorted %s: %.3f ms -> %.3e particles/s" % (method, t[0], n / t[0] * 1e3), flush=True)
        out["paint_tsc_interlaced_cellsorted_%s_ms" % method] = t
    del r2
    c = ComplexField(pm)
    t = timeit(lambda: real.r2c(out=c))
    fb = real.value.numel() * real.value.element_size()
    print("r2c %d^3 %s: %.3f ms  (4x field bytes / t = %.1f GB/s)" % (N, args.dtype, t[0], 4 * fb / t[0] / 1e6), flush=True)
    out["r2c_ms"] = t
    t = timeit(lambda: c.c2r(out=real))
    print("c2r: %.3f ms" % t[0], flush=True)
    out["c2r_ms"] = t
    t = timeit(lambda: c.compensate("CompensateCICShotnoise"))
    print("compensate: %.3f ms (%.1f GB/s)" % (t[0], 2 * fb / t[0] / 1e6), flush=True)
    out["compensate_ms"] = t
    dk = 2 * np.pi / L
    kedges = np.arange(0, np.pi * N / L + dk / 2, dk)
    for Nmu, poles in [(1, []), (5, [0, 2, 4])]:
        mue = np.linspace(-1, 1, Nmu + 1)
        t = timeit(lambda: project_to_basis_device(c, [kedges, mue], poles=poles, is_p3d=False, volume=L ** 3))
        print("power_bin Nmu=%d poles=%s: %.3f ms (%.1f GB/s)" % (Nmu, poles, t[0], fb / t[0] / 1e6), flush=True)
        out["bin_%d_ms" % Nmu] = t
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/microbench_%d_%s.json" % (N, args.dtype), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
