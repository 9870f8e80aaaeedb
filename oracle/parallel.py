"""
TEST INFRASTRUCTURE -- the mesh stages of the reference's FFTPower flow spread over host processes the way the
reference spreads them over MPI ranks: every worker owns a range of x-planes of the complex field, applies the window
compensation (source/mesh/catalog.py:449-594 via Field.apply, one x-slab at a time), forms p3d = c c* V with the
k=0 mode cleared (algorithms/fftpower.py:91-143) and accumulates the per-bin sums of project_to_basis
(:605-672); the parent adds the partial sums (the reference's comm.allreduce, :668-672) and forms the means.

Used by bench.py's CPU arms and by the large-mesh parity tests; never by the product.
Workers are fork()ed so that they see the parent's arrays without a copy.
"""
import multiprocessing as mp
import os

import numpy as np

from . import pmesh_oracle as po

_G = {}


def _worker(span):
    i0, i1 = span
    c, c2 = _G["c"], _G["c2"]
    N, L = _G["N"], _G["L"]
    a = c[i0:i1]
    b = a if c2 is None else c2[i0:i1]
    if _G["comp"] is not None:
        w = po.k_coords(N, L, _G["coord_dtype"], kind="circular")
        wl = [w[0][i0:i1], w[1], w[2]]
        a = po.compensate(_G["comp"], wl, a).astype(c.dtype)
        b = a if c2 is None else po.compensate(_G["comp"], wl, b).astype(c.dtype)
    p3d = a * np.conj(b)                                  # fftpower.py:115-117
    if i0 == 0:
        p3d[0, 0, 0] = 0                                  # :119-124
    p3d = p3d * p3d.dtype.type(L.prod())                  # :128
    x3d = po.k_coords(N, L, _G["coord_dtype"])
    xl = [x3d[0][i0:i1], x3d[1], x3d[2]]
    return po.project_sums(p3d, xl, _G["edges"], los=_G["los"], poles=_G["poles"])


def default_procs():
    return max(1, os.cpu_count() or 1)


def power_from_complex(c1, c2, Nmesh, BoxSize, mode="1d", los=(0, 0, 1), Nmu=5, dk=None, kmin=0., kmax=None,
                       poles=(), coord_dtype="f4", compensation=None, nproc=None, attrs=None):
    """same result dict as pmesh_oracle.power_from_complex (after an optional compensation of the input fields),
    computed by `nproc` processes over x-slabs.  Sums over slabs are added in slab order, so float results may differ
    from the single-process ones in the last bits (as they do between MPI rank counts in the reference)."""
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    L = np.asarray(BoxSize, dtype="f8") * np.ones(3)
    if mode == "1d":
        Nmu = 1
    if dk is None:
        dk = 2 * np.pi / L.min()
    if kmax is None:
        kmax = np.pi * N.min() / L.max() + dk / 2
    kedges = np.arange(kmin, kmax, dk)
    muedges = np.linspace(-1, 1, Nmu + 1, endpoint=True)
    poles = list(poles)
    _poles = [0] + sorted(poles) if 0 not in poles else sorted(poles)
    ell_idx = [_poles.index(l) for l in poles]
    nproc = int(nproc or default_procs())
    nx = c1.shape[0]
    nproc = max(1, min(nproc, nx))
    bounds = np.linspace(0, nx, nproc + 1).astype(int)
    spans = [(int(bounds[i]), int(bounds[i + 1])) for i in range(nproc) if bounds[i + 1] > bounds[i]]
    _G.update(c=c1, c2=c2, N=N, L=L, comp=compensation, coord_dtype=coord_dtype, edges=[kedges, muedges], los=los,
              poles=poles)
    try:
        if len(spans) == 1:
            parts = [_worker(spans[0])]
        else:
            ctx = mp.get_context("fork")
            with ctx.Pool(len(spans)) as pool:
                parts = pool.map(_worker, spans, chunksize=1)
    finally:
        _G.clear()
    xsum = sum(p[0] for p in parts)
    musum = sum(p[1] for p in parts)
    ysum = sum(p[2] for p in parts)
    Nsum = sum(p[3] for p in parts)
    res, pole_res = po.finish_projection(xsum, musum, ysum, Nsum, len(poles) > 0, ell_idx)
    out = dict(kedges=kedges, muedges=muedges, k=res[0], mu=res[1], power=res[2], modes=res[3], attrs=dict(attrs or {}))
    if pole_res is not None:
        out.update(poles_k=pole_res[0], poles_power=pole_res[1], poles_modes=pole_res[2])
    return out
