"""ArrayCatalog -- a catalogue from in-memory arrays (API of nbodykit/source/catalog/array.py:27-86).
Columns may be NumPy arrays, torch tensors on the host, or torch tensors already resident in HBM."""
import numpy
import torch

from ... import CurrentMPIComm
from ...base.catalog import CatalogSource, Column


class ArrayCatalog(CatalogSource):
    """
    data : dict of column name -> array, or a structured NumPy array
    comm : communicator (each rank passes its own share of the rows)
    **kwargs : stored in :attr:`attrs`
    """

    def __repr__(self):
        return "ArrayCatalog(size=%d)" % self.size

    @CurrentMPIComm.enable
    def __init__(self, data, comm=None, **kwargs):
        if not isinstance(data, dict):
            if not (isinstance(data, numpy.ndarray) and data.dtype.names is not None):
                raise ValueError(("input data to ArrayCatalog must have a "
                                  "structured data type with fields"))
            data = {name: data[name] for name in data.dtype.names}
        self.comm = comm
        keys = sorted(data.keys())
        self._source = {k: (v if isinstance(v, torch.Tensor) else numpy.asarray(v)) for k, v in data.items()}
        self._size = len(self._source[keys[0]]) if keys else 0
        for key in keys:
            if len(self._source[key]) != self._size:
                raise ValueError("column `%s` and column `%s` has different size" % (keys[0], key))
        # all ranks must agree on the set of columns
        names = comm.allgather(keys)
        if any(n != names[0] for n in names):
            raise ValueError("mismatch between columns across ranks in ArrayCatalog")
        self.attrs.update(kwargs)
        CatalogSource.__init__(self, comm=comm)

    @property
    def hardcolumns(self):
        defaults = CatalogSource.hardcolumns.fget(self)
        return sorted(set(list(getattr(self, '_source', {}).keys()) + defaults))

    def get_hardcolumn(self, col):
        src = getattr(self, '_source', {})
        if col in src:
            return Column(src[col])
        return CatalogSource.get_hardcolumn(self, col)
