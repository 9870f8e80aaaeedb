"""MultipleSpeciesCatalogMesh -- the summed density of several species on one mesh
(API of nbodykit/source/mesh/species.py:8-182)."""
import logging

from ...utils import attrs_to_dict
from ...pmesh.pm import RealField
from .catalog import CatalogMesh


class MultipleSpeciesCatalogMesh(CatalogMesh):
    logger = logging.getLogger('MultipleSpeciesCatalogMesh')

    def __repr__(self):
        return "(%s as CatalogMesh)" % repr(self.source)

    def __init__(self, source, Nmesh, BoxSize, dtype, selection, position, weight, value, interlaced, compensated,
                 resampler):
        CatalogMesh.__init__(self, source, Nmesh=Nmesh, BoxSize=BoxSize, dtype=dtype, Position=None, Selection=None,
                             Weight=None, Value=None, resampler=resampler, compensated=compensated,
                             interlaced=interlaced)
        self.species = source.species
        self.position = position
        self.selection = selection
        self.weight = weight
        self.value = value

    def __iter__(self):
        return iter(self.species)

    def __getitem__(self, key):
        """the CatalogMesh of one species, with this mesh's parameters"""
        if key not in self.source.species:
            raise KeyError("%s is not a species defined in the source" % key)
        cat = self.source[key]
        mesh = CatalogMesh(cat, BoxSize=self.attrs['BoxSize'], Nmesh=self.attrs['Nmesh'], dtype=self.dtype,
                           Weight=cat[self.weight], Value=cat[self.value], Selection=cat[self.selection],
                           Position=cat[self.position], interlaced=self.interlaced, compensated=self.compensated,
                           resampler=self.resampler)
        return mesh.__finalize__(self)

    def to_complex_field(self, out=None):
        return NotImplemented

    def compute_complex_deferred(self):
        return self.compute(mode='complex'), None

    def to_real_field(self, normalize=True):
        """sum of the species' paints; attrs carry per-species metadata prefixed by the species name and the
        weighted total shot noise  P_shot = sum_i (W_i / W_tot)^2 P_shot,i  (species.py:115-180)"""
        attrs = {'num_per_cell': 0., 'N': 0}
        real = RealField(self.pm)
        real[...] = 0
        for name in self.source.species:
            if self.pm.comm.rank == 0:
                self.logger.info("painting the '%s' species" % name)
            self[name].to_real_field(out=real, normalize=False)
            attrs['num_per_cell'] += real.attrs['num_per_cell']
            attrs['N'] += real.attrs['N']
            attrs.update(attrs_to_dict(real, name + '.'))
        if normalize:
            real /= attrs['num_per_cell']
        real.attrs.clear()
        real.attrs.update(attrs)
        real.attrs['shotnoise'] = 0
        total_weight = sum(real.attrs['%s.W' % name] for name in self.source.species)
        for name in self.source.species:
            real.attrs['shotnoise'] += (real.attrs['%s.W' % name] / total_weight) ** 2 * real.attrs['%s.shotnoise' % name]
        return real
