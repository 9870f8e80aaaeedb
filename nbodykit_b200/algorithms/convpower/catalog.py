"""FKPCatalog -- data + randoms in one catalogue (API of nbodykit/algorithms/convpower/catalog.py:7-259)."""
import logging

import numpy

from ...source.catalog.array import ArrayCatalog
from ...source.catalog.species import MultipleSpeciesCatalog


def FKPWeightFromNbar(P0, nbar):
    """FKP weight 1 / (1 + P0 nbar) for a number-density column (or 1.0 when P0 is None)"""
    if P0 is not None:
        return 1.0 / (1. + P0 * nbar)
    return 1.0


class FKPCatalog(MultipleSpeciesCatalog):
    """
    data, randoms : CatalogSource (randoms may be None: an empty catalogue is used)
    BoxSize : float or 3-vector, optional -- else from the extent of the randoms (+ BoxPad)
    BoxPad : float or 3-vector
    P0 : float, optional -- if given, sets ``FKPWeight = 1/(1 + P0 * nbar)`` on both species
    nbar : str -- name of the n(z) column
    """
    logger = logging.getLogger('FKPCatalog')

    def __repr__(self):
        return "FKPCatalog(species=%s)" % str(self.attrs['species'])

    def __init__(self, data, randoms, BoxSize=None, BoxPad=0.02, P0=None, nbar='NZ'):
        if randoms is None:
            randoms = data[:0]
        MultipleSpeciesCatalog.__init__(self, ['data', 'randoms'], data, randoms)
        for name in self.species:
            if nbar not in self[name]:
                raise ValueError("Column `%s` is not defined in `%s`" % (nbar, name))
        self.nbar = nbar
        if P0 is not None:
            for name in self.species:
                self[name]['FKPWeight'] = FKPWeightFromNbar(P0, self[name][self.nbar])
        else:
            for name in self.species:
                if 'FKPWeight' not in self[name]:
                    self[name]['FKPWeight'] = 1.0
        if numpy.isscalar(BoxSize):
            BoxSize = numpy.ones(3) * BoxSize
        self.attrs['BoxSize'] = BoxSize
        if numpy.isscalar(BoxPad):
            BoxPad = numpy.ones(3) * BoxPad
        self.attrs['BoxPad'] = BoxPad

    def _define_bbox(self, position, selection, species):
        """BoxSize, BoxCenter from the extent of the selected positions of `species` (catalog.py:108-149)"""
        from ...utils import get_data_bounds
        from ...base.catalog import ConstantColumn
        pos, sel = self[species].read([position, selection])
        sel_arr = None
        if not (isinstance(sel, ConstantColumn) and bool(sel.value)):
            sel_arr = sel.compute()
        pos_min, pos_max = get_data_bounds(pos.compute(), self.comm, selection=sel_arr)
        if self.comm.rank == 0:
            self.logger.info("cartesian coordinate range: %s : %s" % (str(pos_min), str(pos_max)))
        if numpy.isinf(pos_min).any() or numpy.isinf(pos_max).any():
            raise ValueError("Range of positions from `%s` is infinite;"
                             "try to use the other species with (bbox_from_species='data'." % species)
        delta = abs(pos_max - pos_min)
        BoxCenter = 0.5 * (pos_min + pos_max)
        if self.attrs['BoxSize'] is None:
            delta *= 1.0 + self.attrs['BoxPad']
            BoxSize = numpy.ceil(delta)
        else:
            BoxSize = self.attrs['BoxSize']
        return BoxSize, BoxCenter

    def to_mesh(self, Nmesh=None, BoxSize=None, BoxCenter=None, dtype='c16', interlaced=False, compensated=False,
                resampler='cic', fkp_weight='FKPWeight', comp_weight='Weight', selection='Selection',
                position='Position', bbox_from_species=None, window=None, nbar=None):
        """mesh that paints the FKP field (convpower/catalog.py:151-259).  dtype 'c16' / 'c8' (default, as in the
        reference) recovers the odd multipoles correctly; 'f8' / 'f4' reproduce the reference's Hermitian short-cut.
        Either way the field is stored Hermitian-compressed on the device (see FKPCatalogMesh)."""
        from .catalogmesh import FKPCatalogMesh
        if window is not None:
            import warnings
            resampler = window
            warnings.warn("the window argument is deprecated. Use resampler= instead", DeprecationWarning)
        for name in self.species:
            for col in [fkp_weight, comp_weight]:
                if col not in self[name]:
                    raise ValueError("the '%s' species is missing the '%s' column" % (name, col))
        if Nmesh is None:
            try:
                Nmesh = self.attrs['Nmesh']
            except KeyError:
                raise ValueError("cannot convert FKP source to a mesh; 'Nmesh' keyword is not "
                                 "supplied and the FKP source does not define one in 'attrs'.")
        if bbox_from_species is not None:
            BoxSize1, BoxCenter1 = self._define_bbox(position, selection, bbox_from_species)
        else:
            if self['randoms'].csize > 0:
                BoxSize1, BoxCenter1 = self._define_bbox(position, selection, "randoms")
            else:
                BoxSize1, BoxCenter1 = self._define_bbox(position, selection, "data")
        if BoxSize is None:
            BoxSize = BoxSize1
        if BoxCenter is None:
            BoxCenter = BoxCenter1
        if self.comm.rank == 0:
            self.logger.info("BoxSize = %s" % str(BoxSize))
            self.logger.info("BoxCenter = %s" % str(BoxCenter))
        return FKPCatalogMesh(self, nbar=self.nbar, comp_weight=comp_weight, fkp_weight=fkp_weight, position=position,
                              value='Value', interlaced=interlaced, compensated=compensated, resampler=resampler,
                              Nmesh=Nmesh, BoxSize=BoxSize, BoxCenter=BoxCenter, dtype=dtype, selection=selection)
