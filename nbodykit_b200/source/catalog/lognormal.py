"""
LogNormalCatalog -- log-normal mock catalogue (API of nbodykit/source/catalog/lognormal.py:53-191), generated ON THE
DEVICE by nbodykit_b200.mockmaker:

  Gaussian delta(k) = white noise * sqrt(P(k)/V), Zel'dovich psi_i(k) = i k_i/k^2 delta(k)     mockmaker.py:83-134
  delta(x), psi(x) by four c2r (this package's CUDA FFT)
  1 + delta_LN = exp(b_L delta)/mean, b_L = bias - 1                                            :213-243, 286
  N_cell ~ Poisson(nbar H^3 (1 + delta_LN)); points at cell node + uniform in-cell offset, emitted in cell order  :300-354
  Position += psi(cell) (periodic wrap), Velocity = f * psi                                     lognormal.py:160-187

Columns `Position`, `Velocity`, `VelocityOffset`: float32 device tensors.  P(k) is tabulated once on the host and
interpolated on the device.  With several ranks every rank builds the same fields (same seed) and keeps the particles of
its own x-slab of generator cells, which is what the reference's generator leaves on each rank: the catalogue does not
depend on the number of ranks.  The realisation is NOT bit-identical to the reference's (pmesh.generate_whitenoise and
mpsort are absent: "parity unpinned"); its statistics are tested (tests/test_gpu_lognormal.py).
"""
import logging
import numbers

import numpy

from ... import CurrentMPIComm, mockmaker
from ...base.catalog import CatalogSource, column
from ...comm import SelfComm
from ...pmesh.pm import ParticleMesh


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
        # the generator mesh lives on this GPU, whatever the size of the catalogue's communicator
        pm = ParticleMesh(BoxSize=self.attrs['BoxSize'], Nmesh=self.attrs['Nmesh'], dtype='f4', comm=SelfComm())
        delta_k, disp_k = mockmaker.gaussian_complex_fields(pm, self.Plin, int(self.attrs['seed']),
                                                            unitary_amplitude=self.attrs['unitary_amplitude'],
                                                            inverted_phase=self.attrs['inverted_phase'],
                                                            compute_displacement=True)
        disp = []
        for d in range(3):
            disp.append(disp_k[0].c2r())
            del disp_k[0]
        delta = delta_k.c2r()
        del delta_k
        # this rank's slab of generator cells (x planes), as the reference's domain decomposition leaves them
        Nx = int(pm.Nmesh[0])
        x0, x1 = (Nx * comm.rank) // comm.size, (Nx * (comm.rank + 1)) // comm.size
        # the Poisson stream is seeded apart from the field's (mockmaker.py:300-306 draws a second seed)
        pos, dsp = mockmaker.poisson_sample_to_points(delta, disp, pm, self.attrs['nbar'], bias=float(self.attrs['bias']) - 1.0,
                                                      seed=(int(self.attrs['seed']) * 2654435761 + 12345) % (2 ** 63),
                                                      x_range=(x0, x1) if comm.size > 1 else None)
        if comm.rank == 0:
            self.logger.info("generated %d particles on a %s mesh" % (int(pos.shape[0]), str([int(v) for v in pm.Nmesh])))
        return pos, dsp
