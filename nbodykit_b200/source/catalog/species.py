"""MultipleSpeciesCatalog -- several catalogues addressed as ``species/column``
(API of nbodykit/source/catalog/species.py:7-262)."""
import logging

import numpy

from ...base.catalog import CatalogSourceBase
from ...utils import attrs_to_dict


class MultipleSpeciesCatalog(CatalogSourceBase):
    """
    names : list of str            species names (column prefix)
    *species : CatalogSource       one catalogue per name (copied)
    """
    logger = logging.getLogger('MultipleSpeciesCatalog')

    def __repr__(self):
        return "MultipleSpeciesCatalog(species=%s)" % str(self.attrs['species'])

    def __init__(self, names, *species, **kwargs):
        if len(set(names)) != len(names):
            raise ValueError("each species must have a unique name")
        if not all(cat.comm is species[0].comm for cat in species):
            raise ValueError("communicator mismatch in MultipleSpeciesCatalog")
        if len(names) != len(species):
            raise ValueError("a name must be provided for each species catalog provided")
        CatalogSourceBase.__init__(self, species[0].comm)
        self.attrs['species'] = list(names)
        for cat, name in zip(species, names):
            self.attrs.update(attrs_to_dict(cat, name + '.'))
        self.attrs.update(kwargs)
        self._sources = {name: cat.copy() for name, cat in zip(names, species)}

    @property
    def size(self):
        return NotImplemented

    @property
    def csize(self):
        return NotImplemented

    def __len__(self):
        raise TypeError("a MultipleSpeciesCatalog has no single length; index a species first")

    @property
    def species(self):
        return self.attrs['species']

    @property
    def columns(self):
        return ['%s/%s' % (sp, col) for sp in self.species for col in self._sources[sp].columns]

    @property
    def hardcolumns(self):
        return ['%s/%s' % (sp, col) for sp in self.species for col in self._sources[sp].hardcolumns]

    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.species:
                return self._sources[key]
            sp, sub = split_column(key, self.species)
            return self._sources[sp][sub]
        raise KeyError("index a MultipleSpeciesCatalog by species name or 'species/column'")

    def __setitem__(self, col, value):
        sp, sub = split_column(col, self.species)
        size = len(self._sources[sp])
        if not numpy.isscalar(value) and len(value) != size:
            raise ValueError("error setting '%s' column, data must be array of size %d, not %d" % (col, size, len(value)))
        self._sources[sp][sub] = value

    def __delitem__(self, col):
        sp, sub = split_column(col, self.species)
        del self._sources[sp][sub]

    def __contains__(self, col):
        return col in self.columns or col in self.species

    def to_mesh(self, Nmesh=None, BoxSize=None, dtype='f4', interlaced=False, compensated=False, resampler='cic',
                weight='Weight', value='Value', selection='Selection', position='Position', window=None):
        from ..mesh.species import MultipleSpeciesCatalogMesh
        if window is not None:
            raise RuntimeError("use resampler instead")
        for name in self.species:
            for col in [position, selection, weight, value]:
                if col not in self[name]:
                    raise ValueError("the '%s' species is missing the '%s' column" % (name, col))
        if BoxSize is None:
            BoxSize = check_species_metadata('BoxSize', self.attrs, self.species)
        if Nmesh is None:
            Nmesh = check_species_metadata('Nmesh', self.attrs, self.species)
        return MultipleSpeciesCatalogMesh(self, Nmesh=Nmesh, BoxSize=BoxSize, dtype=dtype, selection=selection,
                                          position=position, weight=weight, value=value, interlaced=interlaced,
                                          compensated=compensated, resampler=resampler)


def check_species_metadata(name, attrs, species):
    """the single value of ``name`` shared by the catalogue and all species, or an error"""
    vals = []
    if name in attrs:
        vals.append(attrs[name])
    for s in species:
        if s + '.' + name in attrs:
            vals.append(attrs.get(s + '.' + name))
    if len(vals) == 0:
        raise ValueError("please specify ``%s`` attributes" % name)
    if not all(numpy.equal(vals[0], v).all() for v in vals):
        raise ValueError("please specify ``%s`` attributes that are consistent for each species "
                         "and for the multi species catalog; " % name)
    return vals[0]


def split_column(col, species):
    """'species/name' -> (species, name)"""
    fields = col.split('/')
    if len(fields) != 2:
        raise ValueError("new column names should be prefixed by 'species/' where 'species' is one of %s" % str(species))
    sp, sub = fields
    if sp not in species:
        raise ValueError("species '%s' is not valid; should be one of %s" % (sp, str(species)))
    return sp, sub
