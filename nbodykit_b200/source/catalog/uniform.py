"""RandomCatalog / UniformCatalog (API and random streams of nbodykit/source/catalog/uniform.py)."""
import numpy

from ... import CurrentMPIComm
from ...base.catalog import CatalogSource, column
from ...mpirng import MPIRandomState


class RandomCatalog(CatalogSource):
    """a catalogue whose columns can be drawn from `self.rng`, a rank-count-invariant generator"""

    def __repr__(self):
        return "RandomCatalog(size=%d, seed=%s)" % (self.size, self.attrs['seed'])

    @CurrentMPIComm.enable
    def __init__(self, csize, seed=None, comm=None):
        self.comm = comm
        if seed is None:
            if self.comm.rank == 0:
                seed = numpy.random.randint(0, 4294967295)
            seed = self.comm.bcast(seed)
        self.attrs['seed'] = seed
        if csize == 0:
            raise ValueError("no random particles generated!")
        start = comm.rank * csize // comm.size
        end = (comm.rank + 1) * csize // comm.size
        self._size = end - start
        self._rng = MPIRandomState(comm, seed=seed, size=self._size)
        CatalogSource.__init__(self, comm=comm)

    @property
    def rng(self):
        return self._rng


class UniformCatalog(RandomCatalog):
    """uniformly distributed `Position` (in BoxSize) and `Velocity` (in 0.01 BoxSize); the number of
    particles is Poisson(nbar * volume) drawn from `RandomState(seed)` (uniform.py:85-101)"""

    def __repr__(self):
        return "UniformCatalog(size=%d, seed=%s)" % (self.size, self.attrs['seed'])

    @CurrentMPIComm.enable
    def __init__(self, nbar, BoxSize, seed=None, dtype='f8', comm=None):
        self.comm = comm
        _BoxSize = numpy.empty(3, dtype='f8')
        _BoxSize[:] = BoxSize
        self.attrs['BoxSize'] = _BoxSize
        rng = numpy.random.RandomState(seed)
        N = rng.poisson(nbar * numpy.prod(self.attrs['BoxSize']))
        if N == 0:
            raise ValueError("no uniform particles generated, try increasing `nbar` parameter")
        RandomCatalog.__init__(self, N, seed=seed, comm=comm)
        self._pos = (self.rng.uniform(itemshape=(3,)) * self.attrs['BoxSize']).astype(dtype)
        self._vel = (self.rng.uniform(itemshape=(3,)) * self.attrs['BoxSize'] * 0.01).astype(dtype)

    @column
    def Position(self):
        return self.make_column(self._pos)

    @column
    def Velocity(self):
        return self.make_column(self._vel)
