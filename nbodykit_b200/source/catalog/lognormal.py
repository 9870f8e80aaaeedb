"""
LogNormalCatalog -- log-normal mock catalogue (API of nbodykit/source/catalog/lognormal.py:53-191,
algorithm of nbodykit/mockmaker.py:7-359), generated ON THE DEVICE:

  Gaussian delta(k) = white noise * sqrt(P(k)/V)            mockmaker.py:83-124
  Zel'dovich displacement psi_i(k) = i k_i/k^2 delta(k)     :127-134   (3 c2r)
  delta(x) = c2r(delta_k); 1+delta_LN = exp(b_L delta)/mean :213-243   (b_L = bias - 1)
  N_cell ~ Poisson(nbar H^3 (1+delta_LN))                   :300-306
  particles at cell node + uniform in-cell jitter, emitted in cell order  :312-354
  pos += psi(cell) ; wrap                                   lognormal.py:172

Columns `Position`, `Velocity` (= f * psi), `VelocityOffset`: float32 device tensors.  The c2r
transforms are this package's CUDA FFT; random numbers come from torch's Philox generator, so the
realisation is NOT bit-identical to the reference's (which depends on pmesh.generate_whitenoise and
mpsort -- "parity unpinned" in SURVEY.md §2.1); its statistics are (tests/test_gpu_lognormal.py).
"""
import logging
import numbers

import numpy
import torch

from ... import CurrentMPIComm
from ...base.catalog import CatalogSource, column
from ...pmesh.pm import ComplexField, ParticleMesh, RealField


class LogNormalCatalog(CatalogSource):
    logger = logging.getLogger('LogNormalCatalog')

    def __repr__(self):
        return "LogNormalCatalog(seed=%(seed)d, bias=%(bias)g)" % self.attrs

    @CurrentMPIComm.enable
    def __init__(self, Plin, nbar, BoxSize, Nmesh, bias=2., seed=None, cosmo=None, redshift=None,
                 unitary_amplitude=False, inverted_phase=False, growth_rate=None, comm=None):
        self.comm = comm
        self.Plin = Plin
        if cosmo is None:
            cosmo = getattr(self.Plin, 'cosmo', None)
        if redshift is None:
            redshift = getattr(self.Plin, 'redshift', None)
        self.cosmo = cosmo
        self.attrs['nbar'] = nbar
        self.attrs['redshift'] = redshift
        self.attrs['bias'] = bias
        self.attrs['unitary_amplitude'] = unitary_amplitude
        self.attrs['inverted_phase'] = inverted_phase
        _Nmesh = numpy.empty(3, dtype='i8')
        _Nmesh[:] = Nmesh
        self.attrs['Nmesh'] = _Nmesh
        _BoxSize = numpy.empty(3, dtype='f8')
        _BoxSize[:] = BoxSize
        self.attrs['BoxSize'] = _BoxSize
        if seed is None:
            if self.comm.rank == 0:
                seed = numpy.random.randint(0, 4294967295)
            seed = self.comm.bcast(seed)
        if not isinstance(seed, numbers.Integral):
            raise ValueError("the seed used to generate the linear field must be an integer")
        self.attrs['seed'] = seed
        if growth_rate is None:
            growth_rate = 0.0
            if cosmo is not None and hasattr(cosmo, 'scale_independent_growth_rate') and redshift is not None:
                growth_rate = float(cosmo.scale_independent_growth_rate(redshift))
        self.attrs['growth_rate'] = growth_rate
        pos, disp = self._makesource()
        self._pos = pos
        self._vel = disp * growth_rate      # linear velocity in displacement units: v = f * psi
        self._size = int(pos.shape[0])
        CatalogSource.__init__(self, comm=comm)

    @column
    def Position(self):
        return self.make_column(self._pos)

    @column
    def Velocity(self):
        return self.make_column(self._vel)

    @column
    def VelocityOffset(self):
        return self.make_column(self._vel)

    def _makesource(self):
        comm = self.comm
        pm = ParticleMesh(BoxSize=self.attrs['BoxSize'], Nmesh=self.attrs['Nmesh'], dtype='f4', comm=comm)
        if comm.size > 1:
            raise NotImplementedError("multi-rank LogNormalCatalog generation: generate on one rank per slab "
                                      "(see bench.py's per-rank generator) -- the Fourier-space fields here are "
                                      "single-GPU")
        dev = torch.device("cuda", torch.cuda.current_device())
        N = [int(v) for v in pm.Nmesh]
        L = pm.BoxSize
        V = float(L.prod())
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(self.attrs['seed']))
        # white noise with <|w_k|^2> = 1: r2c of unit normals carries 1/N^3 -> scale by sqrt(N^3)
        white = RealField(pm)
        white.value.normal_(generator=gen)
        delta_k = white.r2c()
        del white
        delta_k *= float(numpy.sqrt(numpy.prod(N)))
        if self.attrs['unitary_amplitude']:
            a = delta_k.value.abs()
            delta_k.value /= torch.where(a > 0, a, torch.ones_like(a))
        if self.attrs['inverted_phase']:
            delta_k *= -1.0
        # amplitude sqrt(P(k)/V), evaluated on the host per x-plane with the user's callable
        kx, ky, kz = [numpy.ravel(c).astype('f8') for c in pm.create_coords("complex")]
        amp = torch.empty(delta_k.value.shape, dtype=torch.float32, device=dev)
        kyz2 = ky[:, None] ** 2 + kz[None, :] ** 2
        for i in range(N[0]):
            k = numpy.sqrt(kx[i] ** 2 + kyz2)
            k[k == 0] = 1.0
            p = numpy.asarray(self.Plin(k.ravel())).reshape(k.shape)
            amp[i] = torch.from_numpy(numpy.sqrt(p / V).astype('f4')).to(dev)
        delta_k.value *= amp
        del amp
        delta_k.value[0, 0, 0] = 0
        # Zel'dovich displacement per axis (nearest-grid-point read-out later)
        kxt = torch.from_numpy(kx.astype('f4')).to(dev)[:, None, None]
        kyt = torch.from_numpy(ky.astype('f4')).to(dev)[None, :, None]
        kzt = torch.from_numpy(kz.astype('f4')).to(dev)[None, None, :]
        k2 = kxt ** 2 + kyt ** 2 + kzt ** 2
        k2[0, 0, 0] = 1.0
        disp = []
        for kd in (kxt, kyt, kzt):
            d = ComplexField(pm, (1j * kd / k2) * delta_k.value)
            d.value[0, 0, 0] = 0
            disp.append(d.c2r().value)
            del d
        del k2
        delta = delta_k.c2r()
        del delta_k
        # log-normal transform with the Lagrangian bias (mockmaker.py:213-243, 286)
        bl = float(self.attrs['bias']) - 1.0
        ln = torch.exp(delta.value.double() * bl)
        ln /= ln.mean()
        H3 = float((L / pm.Nmesh).prod())
        lam = ln * (float(self.attrs['nbar']) * H3)
        del ln, delta
        counts = torch.poisson(lam, generator=gen).long().reshape(-1)
        del lam
        ntot = int(counts.sum().item())
        cells = torch.repeat_interleave(torch.arange(counts.numel(), device=dev), counts)   # cell-sorted
        del counts
        iz = cells % N[2]
        iy = (cells // N[2]) % N[1]
        ix = cells // (N[2] * N[1])
        H = [float(L[d] / N[d]) for d in range(3)]
        pos = torch.empty((ntot, 3), dtype=torch.float32, device=dev)
        dsp = torch.empty((ntot, 3), dtype=torch.float32, device=dev)
        for d, idx in enumerate((ix, iy, iz)):
            jitter = torch.rand(ntot, device=dev, dtype=torch.float64, generator=gen)
            dd = disp[d].reshape(-1)[cells]
            x = (idx.double() + jitter) * H[d] + dd.double()
            pos[:, d] = torch.remainder(x, float(L[d])).float()
            dsp[:, d] = dd
        # float32 rounding can land exactly on L: fold it back
        for d in range(3):
            pos[:, d][pos[:, d] >= float(L[d])] = 0.0
        if comm.rank == 0:
            self.logger.info("generated %d particles on a %s mesh" % (ntot, str(N)))
        return pos, dsp
