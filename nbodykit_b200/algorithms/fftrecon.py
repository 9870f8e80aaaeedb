"""
FFTRecon -- FFT based Lagrangian reconstruction in a periodic box (API of nbodykit/algorithms/fftrecon.py:11-269;
schemes LGS / LF2 / LRR of Schmittfull et al. 2015, displacement from the smoothed, bias- and RSD-corrected density).

Every mesh operation runs on the device: tiled / direct window scatter (`pm.paint`), r2c / c2r, the displacement
transfer function `nbk_recon_displacement` (one kernel per direction instead of the reference's Python slab
callback) and the window gather `nbk_readout`.  The result is a MeshSource: `FFTPower(FFTRecon(...), mode='1d')`.
"""
import logging
import warnings

import numpy
import torch

from .._lib import check, darr, lib
from ..base.catalog import CatalogSourceBase
from ..base.mesh import MeshSource
from ..pmesh.pm import ComplexField, ParticleMesh, RealField, _CODE, _ptr, _stream, as_device_tensor


class FFTRecon(MeshSource):
    logger = logging.getLogger('FFTRecon')

    def __repr__(self):
        return "FFTRecon(scheme=%s)" % self.attrs.get('scheme')

    def __init__(self, data, ran, Nmesh, bias=1.0, f=0.0, los=[0, 0, 1], R=20, position='Position',
                 revert_rsd_random=False, scheme='LGS', BoxSize=None):
        assert scheme in ['LGS', 'LF2', 'LRR']
        assert isinstance(data, CatalogSourceBase)
        assert isinstance(ran, CatalogSourceBase)
        comm = data.comm
        assert data.comm == ran.comm
        if Nmesh is None:
            Nmesh = data.attrs['Nmesh']
        _Nmesh = numpy.empty(3, dtype='i8')
        _Nmesh[...] = Nmesh
        if BoxSize is None:
            BoxSize = data.attrs['BoxSize']
        los = numpy.array(los, dtype='f8', copy=True)
        los /= (los ** 2).sum()
        assert len(los) == 3
        assert (~numpy.isnan(los)).all()
        # the reference builds ParticleMesh(BoxSize, Nmesh, comm) -- pmesh's default real type, f8
        pm = ParticleMesh(BoxSize=BoxSize, Nmesh=_Nmesh, dtype='f8', comm=comm)
        if (pm.BoxSize / pm.Nmesh).max() > R:
            if comm.rank == 0:
                warnings.warn("The smoothing radius smaller than the mesh cell size. This may produce undesired numerical results.")
        assert position in data.columns
        assert position in ran.columns
        self.position = position
        MeshSource.__init__(self, comm, pm.Nmesh.copy(), pm.BoxSize.copy(), pm.dtype)
        self.pm = pm
        self.attrs['bias'] = bias
        self.attrs['f'] = f
        self.attrs['los'] = los
        self.attrs['R'] = R
        self.attrs['scheme'] = scheme
        self.attrs['revert_rsd_random'] = bool(revert_rsd_random)
        self.data = data
        self.ran = ran
        if self.comm.rank == 0:
            self.logger.info("Reconstruction for bias=%g, f=%g, smoothing R=%g los=%s" % (bias, f, R, str(los)))
            self.logger.info("Reconstruction scheme = %s" % scheme)

    def to_real_field(self):
        return self.run()

    def run(self):
        s_d, s_r = self._compute_s()
        return self._helper_paint(s_d, s_r)

    # ---- helpers
    def _positions(self, cat):
        """float32 positions on the device (`cat[position].astype('f4')`, fftrecon.py:153,236)"""
        p = cat[self.position].compute()
        t = as_device_tensor(p)
        return t.to(torch.float32).contiguous()

    def work_with(self, cat, s):
        """1 + delta of the (optionally displaced) catalogue: paint(pos - s) / nbar  (fftrecon.py:139-164)"""
        pm = self.pm
        pos = self._positions(cat)
        if s is not None:
            pos = pos - s
        delta = pm.paint(pos, mass=1.0, resampler='cic', hold=False)
        if pm.comm.size > 1:
            # x slabs: the local scatter keeps the planes this rank owns; rows reaching other slabs travel there
            lay = pm.decompose(pos, smoothing=1.0)
            rpos, _ = lay.route(pos, None)
            if rpos.shape[0]:
                pm.paint(rpos, mass=1.0, resampler='cic', hold=True, out=delta)
        nbar = 1.0 * cat.csize / pm.Nmesh.prod()
        delta /= nbar
        return delta

    def _summary_field(self, field, name):
        if self.logger.isEnabledFor(logging.INFO):
            cmean = field.cmean()
            if self.comm.rank == 0:
                self.logger.info("painted %s, mean=%g" % (name, cmean))

    def _helper_paint(self, s_d, s_r):
        """displacements of data and randoms -> the reconstructed density mesh (fftrecon.py:172-211)"""
        def LGS(delta_s_r):
            delta_s_d = self.work_with(self.data, s_d)
            self._summary_field(delta_s_d, "delta_s_d (shifted)")
            delta_s_d -= delta_s_r
            return delta_s_d

        def LRR(delta_s_r):
            delta_s_nr = self.work_with(self.ran, -s_r)
            self._summary_field(delta_s_nr, "delta_s_nr (reverse shifted)")
            delta_d = self.work_with(self.data, None)
            self._summary_field(delta_d, "delta_d (unshifted)")
            delta_s_nr += delta_s_r
            delta_s_nr *= 0.5
            delta_d -= delta_s_nr
            return delta_d

        def LF2(delta_s_r):
            lgs = LGS(delta_s_r)
            lrr = LRR(delta_s_r)
            lgs *= 3.0 / 7.0
            lrr *= 4.0 / 7.0
            lgs += lrr
            return lgs

        delta_s_r = self.work_with(self.ran, s_r)
        self._summary_field(delta_s_r, "delta_s_r (shifted)")
        delta_recon = {'LGS': LGS, 'LF2': LF2, 'LRR': LRR}[self.attrs['scheme']](delta_s_r)
        self._summary_field(delta_recon, "delta_recon")
        return delta_recon

    def _compute_s(self):
        """reconstruction displacements of data and randoms, (n, 3) float32 device tensors (fftrecon.py:213-268)"""
        pm = self.pm
        delta_d = self.work_with(self.data, None)
        self._summary_field(delta_d, "delta_d (unshifted)")
        delta_k = delta_d.r2c()
        code = _CODE[pm.typestr]
        tr, start, count = delta_k._slab()
        los = darr(self.attrs['los'])
        disp = ComplexField(pm)
        # the three displacement fields are transformed once and read out for both catalogues
        fields = []
        for d in range(3):
            check(lib().nbk_recon_displacement(_ptr(delta_k.value), _ptr(disp.value), code, pm._nmesh_c, pm._box_c, tr,
                                               start, count, d, float(self.attrs['R']), float(self.attrs['bias']),
                                               float(self.attrs['f']), los, _stream()), "nbk_recon_displacement")
            fields.append(disp.c2r())

        def solve_displacement(cat):
            dpos = self._positions(cat)
            s = torch.zeros_like(dpos)
            for d in range(3):
                col = fields[d].readout(dpos, resampler='cic', out=torch.empty(dpos.shape[0], dtype=torch.float32,
                                                                             device=dpos.device))
                s[:, d] = col
            return s

        s_d = solve_displacement(self.data)
        if self.logger.isEnabledFor(logging.INFO):
            std = (self.comm.allreduce(float(0) + (s_d.double() ** 2).sum(dim=0).cpu().numpy()) / self.data.csize) ** 0.5
            if self.comm.rank == 0:
                self.logger.info("Solved displacements of data, std(s_d) = %s" % str(std))
        s_r = solve_displacement(self.ran)
        if self.logger.isEnabledFor(logging.INFO):
            std = (self.comm.allreduce(float(0) + (s_r.double() ** 2).sum(dim=0).cpu().numpy()) / self.ran.csize) ** 0.5
            if self.comm.rank == 0:
                self.logger.info("Solved displacements of randoms, std(s_r) = %s" % str(std))
        # convention 1: the data displacement also reverts RSD; convention 2: the randoms' as well (fftrecon.py:260-266)
        factor = torch.tensor(1 + self.attrs['los'] * self.attrs['f'], dtype=torch.float64, device=s_d.device)
        s_d = (s_d.double() * factor).float()
        if self.attrs['revert_rsd_random']:
            s_r = (s_r.double() * factor).float()
        return s_d, s_r
