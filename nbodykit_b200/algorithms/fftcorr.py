"""
FFTCorr -- correlation function xi(r), xi(r, mu), xi_l(r) of a periodic box through FFTs
(API of nbodykit/algorithms/fftcorr.py:9-235): the 3-D power c1 c2* V (zero mode cleared) is transformed back
with c2r, divided by V, and binned in the wrapped separation with the same kernel that bins P(k)
(`nbk_power_bin`, real input, coordinates index * L/N, mu edges linspace(0, 1)).
"""
import logging

import numpy

from .._lib import check, lib
from ..binned_statistic import BinnedStatistic
from ..pmesh.pm import ComplexField, _CODE, _ptr, _stream
from .fftpower import _find_unique_edges, FFTBase, project_to_basis_device


class FFTCorr(FFTBase):
    """
    first, mode ('1d' | '2d'), Nmesh, BoxSize, second, los, Nmu, dr, rmin, rmax, poles -- as the reference.
    Results: `.corr` (BinnedStatistic over 'r' [, 'mu']: `r`, [`mu`,] `corr`, `modes`), `.poles` (`corr_L`), `.attrs`.
    """
    logger = logging.getLogger('FFTCorr')

    def __init__(self, first, mode, Nmesh=None, BoxSize=None, second=None, los=[0, 0, 1], Nmu=5, dr=None, rmin=0.,
                 rmax=None, poles=[]):
        if mode not in ['1d', '2d']:
            raise ValueError("`mode` should be either '1d' or '2d'")
        if poles is None:
            poles = []
        if numpy.isscalar(los) or len(los) != 3:
            raise ValueError("line-of-sight ``los`` should be vector with length 3")
        if not numpy.allclose(numpy.einsum('i,i', los, los), 1.0, rtol=1e-5):
            raise ValueError("line-of-sight ``los`` must be a unit vector")
        FFTBase.__init__(self, first, second, Nmesh, BoxSize)
        self.attrs['mode'] = mode
        self.attrs['los'] = los
        self.attrs['Nmu'] = Nmu
        self.attrs['poles'] = poles
        if dr is None:
            dr = self.attrs['BoxSize'].min() / self.attrs['Nmesh'].max()
        self.attrs['dr'] = dr
        self.attrs['rmin'] = rmin
        self.attrs['rmax'] = rmax
        self.corr, self.poles = self.run()
        self.attrs.update(self.corr.attrs)

    def run(self):
        if self.attrs['mode'] == "1d":
            self.attrs['Nmu'] = 1
        c1, c2, attrs = self._compute_3d_power(self.first, self.second)
        comp1, comp2 = self._deferred_compensation
        if comp1:
            c1.compensate(comp1)
        if comp2 and c2 is not c1:
            c2.compensate(comp2)
        pm = c1.pm
        if c2 is not c1:
            if tuple(c2.value.shape) != tuple(c1.value.shape):
                raise ValueError("FFTCorr: mesh shape mismatch between the two sources")
            if c2.pm.typestr != pm.typestr:      # NumPy promotion of c1 * conj(c2), as in the reference
                if pm.typestr == 'f8':
                    c2 = ComplexField(pm, c2.value.to(c1.value.dtype))
                else:
                    c1 = ComplexField(c2.pm, c1.value.to(c2.value.dtype))
                    pm = c1.pm
        V = float(pm.BoxSize.prod())
        # 3-D power with the zero mode cleared; xi = c2r(P3d) / V
        p3d = ComplexField(pm)
        owns_zero = (pm.y_start == 0) if pm.transposed else True
        check(lib().nbk_cross_power(_ptr(c1.value), _ptr(c2.value) if c2 is not c1 else None, _ptr(p3d.value),
                                    _CODE[pm.typestr], p3d.value.numel(), V, 1 if owns_zero else 0, _stream()),
              "nbk_cross_power")
        y3d = p3d.c2r()
        y3d *= 1.0 / V
        dr, rmin, rmax = self.attrs['dr'], self.attrs['rmin'], self.attrs['rmax']
        if rmax is None:
            rmax = 0.5 * pm.BoxSize.min() + dr / 2
        if dr > 0:
            redges = numpy.arange(rmin, rmax, dr)
            rcenters = None
        else:
            redges, rcenters = _find_unique_edges(pm, rmax, real=True)
        muedges = numpy.linspace(0, 1, self.attrs['Nmu'] + 1, endpoint=True)
        edges = [redges, muedges]
        coords = [rcenters, None]
        result, pole_result = project_to_basis_device(y3d, edges, poles=self.attrs['poles'], los=self.attrs['los'],
                                                      is_p3d=True)
        # xi is a RealField: the reference bins it into sums of the field's real dtype (fftpower.py:583, 'corr' is f8)
        result = list(result)
        result[2] = numpy.ascontiguousarray(result[2].real)
        if pole_result is not None:
            pole_result = (pole_result[0], numpy.ascontiguousarray(pole_result[1].real), pole_result[2])
        if self.attrs['mode'] == "1d":
            cols, icols = ['r', 'corr', 'modes'], [0, 2, 3]
            edges, coords = edges[0:1], coords[0:1]
        else:
            cols, icols = ['r', 'mu', 'corr', 'modes'], [0, 1, 2, 3]
        dtype = numpy.dtype([(name, result[icol].dtype.str) for icol, name in zip(icols, cols)])
        corr = numpy.squeeze(numpy.empty(result[0].shape, dtype=dtype))
        for icol, col in zip(icols, cols):
            corr[col][:] = numpy.squeeze(result[icol])
        poles = None
        if pole_result is not None:
            r, poles_, N = pole_result
            cols = ['r'] + ['corr_%d' % l for l in self.attrs['poles']] + ['modes']
            res = [r] + [pole for pole in poles_] + [N]
            dtype = numpy.dtype([(name, res[icol].dtype.str) for icol, name in enumerate(cols)])
            poles = numpy.empty(res[0].shape, dtype=dtype)
            for icol, col in enumerate(cols):
                poles[col][:] = res[icol]
        return self._make_datasets(edges, poles, corr, coords, attrs)

    def __getstate__(self):
        return dict(corr=self.corr.__getstate__(), poles=self.poles.__getstate__() if self.poles is not None else None,
                    attrs=self.attrs)

    def __setstate__(self, state):
        self.attrs = state['attrs']
        self.corr = BinnedStatistic.from_state(state['corr'])
        self.poles = None
        if state['poles'] is not None:
            self.poles = BinnedStatistic.from_state(state['poles'])

    def _make_datasets(self, edges, poles, corr, coords, attrs):
        if self.attrs['mode'] == '1d':
            corr = BinnedStatistic(['r'], edges, corr, fields_to_sum=['modes'], coords=coords, **attrs)
        else:
            corr = BinnedStatistic(['r', 'mu'], edges, corr, fields_to_sum=['modes'], coords=coords, **attrs)
        if poles is not None:
            poles = BinnedStatistic(['r'], [corr.edges['r']], poles, fields_to_sum=['modes'],
                                    coords=[corr.coords['r']], **attrs)
        return corr, poles
