"""
Device-side mock makers behind LogNormalCatalog and LinearMesh (what nbodykit/mockmaker.py:7-359 computes, laid out
for one GPU holding the whole generator mesh in HBM):

  gaussian_complex_fields   delta(k) = white noise * sqrt(P(k)/V) [+ Zel'dovich psi_i(k) = i k_i/k^2 delta(k)]  (:7-134)
  lognormal_transform       1 + delta_LN = exp(b_L delta) / <exp(b_L delta)>                                      (:213-243)
  poisson_sample_to_points  N_cell ~ Poisson(nbar H^3 (1 + delta_LN)); points at cell node + uniform in-cell
                            offset, emitted in cell order, displaced by psi(cell)                                 (:245-359)

The user's P(k) callable is evaluated ONCE on the host, on a logarithmic wavenumber grid covering the mesh, and
interpolated on the device (log-log linear) -- the per-x-plane host evaluation + H2D copy of the earlier version is gone.
Random numbers come from torch's counter-based Philox generator on the full generator mesh: a catalogue built on P
ranks generates the same field on every rank (180 GB of HBM hold a 1024^3 generator mesh many times over) and keeps its
own slab of cells, so the realisation does not depend on the number of ranks.  It is NOT bit-identical to the
reference's (pmesh.generate_whitenoise + mpsort are absent from /root/reference: "parity unpinned");
its statistics are tested (tests/test_gpu_lognormal.py).
"""
import numbers

import numpy
import torch

from .pmesh.pm import ComplexField, RealField

PLANES = 16           # x-planes of the generator mesh handled per batch of elementwise work


class PowerTable(object):
    """P(k) tabulated on a log grid (host evaluation of the callable, once) and interpolated on the device"""

    def __init__(self, linear_power, kmin, kmax, device, n=8192):
        lk = numpy.linspace(numpy.log(kmin * 0.999), numpy.log(kmax * 1.001), n)
        p = numpy.asarray(linear_power(numpy.exp(lk)), dtype='f8').reshape(-1)
        if p.shape[0] != n:
            raise ValueError("the linear power callable must return one value per wavenumber")
        self.positive = bool((p > 0).all())
        self.lk0 = float(lk[0])
        self.inv_dlk = float((n - 1) / (lk[-1] - lk[0]))
        self.n = n
        self.tab = torch.from_numpy(numpy.log(p) if self.positive else p).to(device)

    def __call__(self, k):
        """k: device tensor of wavenumbers > 0 (any shape) -> P(k), float64"""
        t = (torch.log(k.double()) - self.lk0) * self.inv_dlk
        t = t.clamp_(0.0, self.n - 1 - 1e-9)
        i = t.floor().long()
        w = t - i.double()
        v = self.tab[i] * (1.0 - w) + self.tab[i + 1] * w
        return torch.exp(v) if self.positive else v


def _k_axes(pm, device):
    """float64 wavenumber vectors of the (single-GPU, untransposed) complex layout"""
    ks = []
    for d in range(3):
        n = int(pm.Nmesh[d]) if d < 2 else int(pm.Nmesh[2]) // 2 + 1
        j = torch.arange(n, device=device, dtype=torch.float64)
        N = int(pm.Nmesh[d])
        j = torch.where(j >= (N + 1) // 2, j - N, j)
        ks.append(j * (2 * numpy.pi / float(pm.BoxSize[d])))
    return ks


def gaussian_complex_fields(pm, linear_power, seed, unitary_amplitude=False, inverted_phase=False,
                            compute_displacement=False, logger=None):
    """Gaussian delta(k) with power spectrum `linear_power` on the single-GPU mesh `pm`, optionally with the three
    Zel'dovich displacement fields psi_i(k) = i k_i / k^2 delta(k).  Returns (delta_k, [psi_x, psi_y, psi_z] | None) as
    ComplexFields (mockmaker.py:7-134; normalisation: <|delta_k|^2> = P(k)/V, the zero mode is cleared)."""
    if not isinstance(seed, numbers.Integral):
        raise ValueError("the seed used to generate the linear field must be an integer")
    if pm.comm.size != 1:
        raise ValueError("the generator mesh lives on one GPU (every rank builds the same field)")
    dev = torch.device("cuda", torch.cuda.current_device())
    N = [int(v) for v in pm.Nmesh]
    V = float(pm.BoxSize.prod())
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    # white noise with <|w_k|^2> = 1: r2c of unit normals carries 1/N^3 -> scale by sqrt(N^3)
    white = RealField(pm)
    white.value.normal_(generator=gen)
    delta_k = white.r2c()
    del white
    amp0 = float(numpy.sqrt(numpy.prod(N)))
    if inverted_phase:
        amp0 = -amp0
    kx, ky, kz = _k_axes(pm, dev)
    kfund = 2 * numpy.pi / float(pm.BoxSize.max())
    knyq = numpy.sqrt(sum((numpy.pi * N[d] / float(pm.BoxSize[d])) ** 2 for d in range(3)))
    table = PowerTable(linear_power, kfund * 0.5, knyq * 1.01, dev)
    disp = [ComplexField(pm) for _ in range(3)] if compute_displacement else None
    kyz2 = ky[:, None] ** 2 + kz[None, :] ** 2
    cdt = delta_k.value.dtype
    for x0 in range(0, N[0], PLANES):
        x1 = min(N[0], x0 + PLANES)
        k2 = kx[x0:x1, None, None] ** 2 + kyz2[None]
        zero = k2 == 0
        k2 = torch.where(zero, torch.ones_like(k2), k2)
        amp = torch.sqrt(table(torch.sqrt(k2)) / V)
        blk = delta_k.value[x0:x1]
        if unitary_amplitude:
            a = blk.abs()
            blk /= torch.where(a > 0, a, torch.ones_like(a))
            blk *= (amp if not inverted_phase else -amp).to(blk.real.dtype)
        else:
            blk *= (amp * amp0).to(blk.real.dtype)
        blk[zero] = 0
        if compute_displacement:
            f = (1j * blk.to(torch.complex128)) / k2
            disp[0].value[x0:x1] = (f * kx[x0:x1, None, None]).to(cdt)
            disp[1].value[x0:x1] = (f * ky[None, :, None]).to(cdt)
            disp[2].value[x0:x1] = (f * kz[None, None, :]).to(cdt)
            for d in range(3):
                disp[d].value[x0:x1][zero] = 0
    if logger is not None:
        logger.info("gaussian field generated on a %s mesh" % str(N))
    return delta_k, disp


def lognormal_transform(density, bias=1.):
    """in place: 1 + delta_LN = exp(bias * delta) / mean(exp(bias * delta))  (mockmaker.py:213-243); returns the field"""
    v = density.value
    v.mul_(float(bias)).exp_()
    v.div_(v.mean(dtype=torch.float64).to(v.dtype))
    return density


def poisson_sample_to_points(delta, displacement, pm, nbar, bias=1., seed=None, logger=None, x_range=None):
    """Poisson-sample the log-normal transform of `delta` (RealField) to points, displaced by `displacement`
    (3 RealFields, nearest-grid-point).  Returns (pos, disp): float32 (n, 3) device tensors in CELL ORDER (x slowest).
    x_range = (x0, x1): keep only the points of the cell planes [x0, x1) -- the share of one rank of a catalogue that
    several ranks generate identically."""
    dev = delta.value.device
    N = [int(v) for v in pm.Nmesh]
    L = [float(v) for v in pm.BoxSize]
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) if seed is not None else 0)
    lognormal_transform(delta, bias)
    H3 = float(numpy.prod([L[d] / N[d] for d in range(3)]))
    lam = delta.value.double().mul_(float(nbar) * H3)
    counts = torch.poisson(lam, generator=gen).to(torch.int32)
    del lam
    ntot_all = int(counts.sum(dtype=torch.int64).item())
    plane = N[1] * N[2]
    x0, x1 = (0, N[0]) if x_range is None else (int(x_range[0]), int(x_range[1]))
    # the in-cell offsets of ALL points are drawn (three streams of the global length), then the slab is cut out: the
    # points of a cell do not depend on who keeps them
    cnt_flat = counts.reshape(-1)
    before = int(cnt_flat[:x0 * plane].sum(dtype=torch.int64).item()) if x0 > 0 else 0
    mine = cnt_flat[x0 * plane:x1 * plane]
    n = int(mine.sum(dtype=torch.int64).item())
    cells = torch.repeat_interleave(torch.arange(x0 * plane, x1 * plane, device=dev), mine.long())   # cell-sorted
    del counts, cnt_flat, mine
    H = [L[d] / N[d] for d in range(3)]
    pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
    dsp = torch.empty((n, 3), dtype=torch.float32, device=dev)
    for d in range(3):
        jitter = torch.rand(ntot_all, device=dev, dtype=torch.float32, generator=gen)[before:before + n].double()
        dd = displacement[d].value.reshape(-1)[cells] if displacement is not None else torch.zeros(n, device=dev)
        index = (cells // plane) if d == 0 else (((cells // N[2]) % N[1]) if d == 1 else (cells % N[2]))
        x = (index.double() + jitter) * H[d] + dd.double()
        del index
        del jitter
        p = torch.remainder(x, L[d]).float()
        p[p >= L[d]] = 0.0          # float32 rounding can land exactly on L: fold it back
        pos[:, d] = p
        dsp[:, d] = dd
    if logger is not None:
        logger.info("poisson sampling done: %d of %d points kept" % (n, ntot_all))
    return pos, dsp
