"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's FFTRecon flow and of the pmesh readout it relies on.
Nothing under nbodykit_b200/ may import this module.

Follows /root/reference/nbodykit/algorithms/fftrecon.py:
  work_with            :139-164   paint(f4 positions [- s]) / nbar with the pm default window (cic)
  kernel / _compute_s  :213-268   displacement modes  i k_d/k^2 * delta_k * exp(-k^2 R^2/2) / (b (1 + f/b mu^2)),
                                  c2r, readout at the particle positions, RSD factor (1 + los f)
  _helper_paint        :172-211   LGS / LRR / LF2 combinations
The pmesh pieces (paint, readout, r2c/c2r conventions) are the restatements of oracle/pmesh_oracle.py; pmesh itself is
not vendored in /root/reference, so this flow is "parity unpinned" against the reference binary.
"""
import numpy as np

from . import pmesh_oracle as po


def readout(field, pos, Nmesh, BoxSize, resampler="cic", shift=0.0):
    """RealField.readout: sum over the stencil of W * field[cell], f8 accumulation (transpose of po.paint)"""
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8"))
    pos = np.asarray(pos)
    out = np.zeros(len(pos), dtype="f8")
    if len(pos) == 0:
        return out
    g = po.grid_coords(pos, N, BoxSize, shift, 0)
    i0, w = [], []
    for d in range(3):
        a, b = po.window_1d(g[:, d], resampler)
        i0.append(a)
        w.append(b)
    sup = po.SUPPORT[resampler]
    f = np.asarray(field, dtype="f8")
    for rx in range(sup):
        ix = (i0[0] + rx) % N[0]
        for ry in range(sup):
            iy = (i0[1] + ry) % N[1]
            wxy = w[0][rx] * w[1][ry]
            for rz in range(sup):
                iz = (i0[2] + rz) % N[2]
                out += (wxy * w[2][rz]) * f[ix, iy, iz]
    return out


def displacement_modes(delta_k, Nmesh, BoxSize, axis, R, bias, f, los):
    """the `kernel(d)` callback of fftrecon.py:215-230 on the Hermitian-compressed field, f8 wavenumbers"""
    k = po.k_coords(Nmesh, BoxSize, coord_dtype="f8", kind="wavenumber")
    k2 = sum(ki ** 2 for ki in k)
    k2 = np.where(k2 == 0, 1.0, k2)
    mu = sum(k[i] * los[i] for i in range(3)) / k2 ** 0.5
    v = delta_k * np.exp(-0.5 * k2 * R ** 2)
    v = v / (bias * (1 + f / bias * mu ** 2))
    return 1j * k[axis] / k2 * v


def work_with(pos, s, ncat, Nmesh, BoxSize):
    p = np.asarray(pos, dtype="f4")
    if s is not None:
        p = p - s
    nbar = 1.0 * ncat / np.prod(np.asarray(Nmesh, dtype="f8") * np.ones(3))
    return po.paint(p, None, Nmesh, BoxSize, "cic", dtype="f8") / nbar


def fftrecon(data, ran, Nmesh, BoxSize, bias=1.0, f=0.0, los=(0, 0, 1), R=20., scheme="LGS", revert_rsd_random=False):
    """reconstructed density mesh (f8) and the two displacement arrays (f4)"""
    los = np.array(los, dtype="f8")
    los /= (los ** 2).sum()
    N3 = np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8")
    delta_k = po.r2c(work_with(data, None, len(data), Nmesh, BoxSize))
    fields = [po.c2r(displacement_modes(delta_k, Nmesh, BoxSize, d, R, bias, f, los), N3) for d in range(3)]

    def solve(cat):
        dpos = np.asarray(cat, dtype="f4")
        s = np.zeros_like(dpos, dtype="f4")
        for d in range(3):
            s[:, d] = readout(fields[d], dpos, Nmesh, BoxSize, "cic")
        return s
    s_d, s_r = solve(data), solve(ran)
    s_d = (s_d * (1 + los * f)).astype("f4")
    if revert_rsd_random:
        s_r = (s_r * (1 + los * f)).astype("f4")

    delta_s_r = work_with(ran, s_r, len(ran), Nmesh, BoxSize)

    def LGS():
        return work_with(data, s_d, len(data), Nmesh, BoxSize) - delta_s_r

    def LRR():
        delta_s_nr = work_with(ran, -s_r, len(ran), Nmesh, BoxSize)
        delta_d = work_with(data, None, len(data), Nmesh, BoxSize)
        return delta_d - 0.5 * (delta_s_nr + delta_s_r)
    if scheme == "LGS":
        out = LGS()
    elif scheme == "LRR":
        out = LRR()
    else:
        out = LGS() * (3.0 / 7.0) + LRR() * (4.0 / 7.0)
    return out, s_d, s_r
