"""
TEST INFRASTRUCTURE -- NumPy restatement of ConvolvedFFTPower (nbodykit/algorithms/convpower/fkp.py:408-655 with
catalogmesh.py:122-244 and catalog.py:108-149).  Never imported by the product.

PARITY STATUS: the reference's ConvolvedFFTPower cannot run here (pmesh absent) -> **parity unpinned** for the
multipole values; pinned pieces: the real Y_lm polynomials (golden values from the reference's own
get_real_Ylm, tests/golden/ylm_reference.npz), project_to_basis / Compensate* (shared with the FFTPower oracle),
and the reference's own identities re-run on both oracle and GPU (data.norm = N nbar style checks,
shot-noise identity, algorithms/tests/test_conv_power.py:173-180).
"""
import json
import os

import numpy as np

from . import pmesh_oracle as po

_TABLE = None


def ylm_table():
    global _TABLE
    if _TABLE is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ylm_table.json")
        _TABLE = json.load(open(path))["terms"]
    return _TABLE


def real_ylm(l, m, x, y, z):
    out = 0.0
    for c, px, py, pz in ylm_table()["%d,%d" % (l, m)]:
        out = out + c * x ** px * y ** py * z ** pz
    return out


def x_coords(Nmesh, BoxSize, coord_dtype="f4"):
    """real-space field coordinates, wrapped to [-L/2, L/2): x = fl32(f32(i - N [i >= N/2]) * f32(H))"""
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    L = np.asarray(BoxSize, dtype="f8") * np.ones(3)
    ct = np.dtype(coord_dtype).type
    out = []
    for d in range(3):
        i = po.freq_index(int(N[d]))
        x = i.astype(ct) * ct(L[d] / N[d])
        shape = [1, 1, 1]
        shape[d] = len(x)
        out.append(x.reshape(shape))
    return out


def fkp_field(dpos, rpos, wd, wr, Nmesh, BoxSize, BoxCenter, resampler="cic", dtype="f8"):
    """F = (paint(data; w) - alpha paint(randoms; w)) / V_cell, alpha given by the caller through wd/wr sums of
    COMPLETENESS weights -- here wd/wr are (completeness, fkp) pairs"""
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    L = np.asarray(BoxSize, dtype="f8") * np.ones(3)
    C = np.asarray(BoxCenter, dtype="f8") * np.ones(3)
    (wcd, wfd), (wcr, wfr) = wd, wr
    alpha = wcd.sum() / wcr.sum()
    fd = po.paint(dpos - C, wcd * wfd, N, L, resampler)
    fr = po.paint(rpos - C, wcr * wfr, N, L, resampler)
    F = (fd - alpha * fr) / (L / N).prod()
    return F.astype(dtype), alpha


def convpower(dpos, rpos, wd, wr, nbar_d, nbar_r, Nmesh, BoxSize, BoxCenter, poles, resampler="cic", dtype="f8",
              dk=None, kmin=0., coord_dtype="f4"):
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    L = np.asarray(BoxSize, dtype="f8") * np.ones(3)
    C = np.asarray(BoxCenter, dtype="f8") * np.ones(3)
    F, alpha = fkp_field(dpos, rpos, wd, wr, N, L, C, resampler, dtype)
    V = L.prod()
    comp = po.COMPENSATION.get((False, resampler))
    wc = po.k_coords(N, L, coord_dtype, kind="circular")
    cfield = po.r2c(F)
    if comp is not None:
        cfield = po.compensate(comp, wc, cfield).astype(cfield.dtype)
    A0 = cfield * V
    (wcd, wfd), (wcr, wfr) = wd, wr
    norm_d = float((nbar_d * wcd * wfd * wfd).sum())
    norm_r = float((nbar_r * wcr * wfr * wfr).sum()) * alpha
    norm = 1.0 / norm_r
    shot = (float((wcd ** 2 * wfd * wfd).sum()) + alpha ** 2 * float((wcr ** 2 * wfr * wfr).sum())) / norm_r
    if dk is None:
        dk = 2 * np.pi / L.min()
    kmax = np.pi * N.min() / L.max() + dk / 2
    kedges = np.arange(kmin, kmax, dk)
    edges = [kedges, np.linspace(-1, 1, 2)]
    kx3 = po.k_coords(N, L, coord_dtype)
    offset = C + 0.5 * L / N
    xg = [x.astype("f8") + offset[i] for i, x in enumerate(x_coords(N, L, coord_dtype))]
    xnorm = np.sqrt(sum(x ** 2 for x in xg))
    xg = [x / xnorm for x in xg]
    kg = [k.astype("f8") for k in kx3]
    knorm = np.sqrt(sum(k ** 2 for k in kg))
    knorm[knorm == 0.] = np.inf
    kg = [k / knorm for k in kg]
    out = dict(alpha=alpha, data_norm=norm_d, randoms_norm=norm_r, shotnoise=shot, kedges=kedges)
    ells = sorted(poles)
    for ell in [l for l in ells if l > 0]:
        Aell = np.zeros_like(A0)
        for m in range(-ell, ell + 1):
            r = (F * real_ylm(ell, m, xg[0], xg[1], xg[2])).astype(F.dtype)
            c = po.r2c(r)
            Aell += c * real_ylm(ell, m, kg[0], kg[1], kg[2])
        if comp is not None:
            Aell = po.compensate(comp, wc, Aell).astype(Aell.dtype)
        Aell = Aell * (4 * np.pi * V)
        P = norm * A0 * np.conj(Aell)
        res, _ = po.project_to_basis(P, kx3, edges)
        out["power_%d" % ell] = np.squeeze(res[2])
    P0 = norm * A0 * np.conj(A0)
    res, _ = po.project_to_basis(P0, kx3, edges)
    out["power_0"] = np.squeeze(res[2])
    out["k"] = np.squeeze(res[0])
    out["modes"] = np.squeeze(res[3])
    return out


def convpower_full(dpos, rpos, wd, wr, nbar_d, nbar_r, Nmesh, BoxSize, BoxCenter, poles, resampler="cic",
                   dk=None, kmin=0., coord_dtype="f4"):
    """ConvolvedFFTPower on a FULL complex mesh (the reference's default dtype='c16', convpower/catalog.py:169-176):
    complex-to-complex transforms of the N^3 field and project_to_basis over every mode (hermitian_symmetric=False,
    fftpower.py:572).  Restated with numpy.fft.fftn; it is what pins the odd multipoles of the GPU path, which stores
    the same fields Hermitian-compressed and folds the mirror half with the anti-Hermitian sign."""
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    L = np.asarray(BoxSize, dtype="f8") * np.ones(3)
    C = np.asarray(BoxCenter, dtype="f8") * np.ones(3)
    F, alpha = fkp_field(dpos, rpos, wd, wr, N, L, C, resampler, "f8")
    V = L.prod()
    ct = np.dtype(coord_dtype).type
    # full (uncompressed) coordinate arrays on all three axes
    kx3, wc = [], []
    for d in range(3):
        j = po.freq_index(int(N[d]))
        shape = [1, 1, 1]
        shape[d] = int(N[d])
        kx3.append((j.astype(ct) * ct(2 * np.pi / L[d])).reshape(shape))
        wc.append((j.astype(ct) * ct(2 * np.pi / N[d])).reshape(shape))
    comp = po.COMPENSATION.get((False, resampler))

    def c2c(x):
        return np.fft.fftn(x) / float(np.prod(N))
    cfield = c2c(F)
    if comp is not None:
        cfield = po.compensate(comp, wc, cfield)
    A0 = cfield * V
    (wcd, wfd), (wcr, wfr) = wd, wr
    norm_r = float((nbar_r * wcr * wfr * wfr).sum()) * alpha
    norm = 1.0 / norm_r
    if dk is None:
        dk = 2 * np.pi / L.min()
    kmax = np.pi * N.min() / L.max() + dk / 2
    kedges = np.arange(kmin, kmax, dk)
    edges = [kedges, np.linspace(-1, 1, 2)]
    offset = C + 0.5 * L / N
    xg = [x.astype("f8") + offset[i] for i, x in enumerate(x_coords(N, L, coord_dtype))]
    xnorm = np.sqrt(sum(x ** 2 for x in xg))
    xg = [x / xnorm for x in xg]
    kg = [k.astype("f8") for k in kx3]
    knorm = np.sqrt(sum(k ** 2 for k in kg))
    knorm[knorm == 0.] = np.inf
    kg = [k / knorm for k in kg]
    out = dict(alpha=alpha, kedges=kedges)
    for ell in sorted(poles):
        if ell == 0:
            P = norm * A0 * np.conj(A0)
        else:
            Aell = np.zeros_like(A0)
            for m in range(-ell, ell + 1):
                Aell += c2c(F * real_ylm(ell, m, xg[0], xg[1], xg[2])) * real_ylm(ell, m, kg[0], kg[1], kg[2])
            if comp is not None:
                Aell = po.compensate(comp, wc, Aell)
            P = norm * A0 * np.conj(Aell * (4 * np.pi * V))
        res, _ = po.project_to_basis(P, kx3, edges, hermitian_symmetric=False)
        out["power_%d" % ell] = np.squeeze(res[2])
        out["k"] = np.squeeze(res[0])
        out["modes"] = np.squeeze(res[3])
    return out
