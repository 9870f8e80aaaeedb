"""
RandomCatalog / UniformCatalog -- catalogues drawn from a generator whose stream does not depend on the number of
ranks (API of nbodykit/source/catalog/uniform.py:8-115).

What is pinned by the reference and therefore fixed here: the particle count is `RandomState(seed).poisson(nbar V)`,
rank r owns the rows [r N / P, (r + 1) N / P), and `Position` then `Velocity` are the first two `.uniform(itemshape=(3,))`
draws of `MPIRandomState(seed)` scaled by BoxSize and 0.01 BoxSize -- which makes BASELINE config 1 reproducible bit
for bit (tests/test_oracle_vs_reference.py).  The columns are host arrays; `to_mesh()` moves them to the device once.
"""
import numpy

from ... import CurrentMPIComm
from ...base.catalog import CatalogSource, column
from ...mpirng import MPIRandomState


def _agree_on_seed(comm, seed):
    """one seed for all ranks (drawn on rank 0 when the caller gave none)"""
    if seed is None:
        seed = comm.bcast(numpy.random.randint(0, 4294967295) if comm.rank == 0 else None)
    return seed


def _row_share(comm, csize):
    """number of rows of rank r: the split [r N / P, (r + 1) N / P) the reference uses"""
    return (comm.rank + 1) * csize // comm.size - comm.rank * csize // comm.size


class RandomCatalog(CatalogSource):
    """`csize` rows spread over the ranks, with `self.rng` to draw further columns from"""

    def __repr__(self):
        return "RandomCatalog(size=%d, seed=%s)" % (self.size, self.attrs['seed'])

    @CurrentMPIComm.enable
    def __init__(self, csize, seed=None, comm=None):
        if csize == 0:
            raise ValueError("no random particles generated!")
        self.comm = comm
        self.attrs['seed'] = _agree_on_seed(comm, seed)
        self._size = _row_share(comm, csize)
        self._rng = MPIRandomState(comm, seed=self.attrs['seed'], size=self._size)
        CatalogSource.__init__(self, comm=comm)

    @property
    def rng(self):
        """the :class:`MPIRandomState` of this catalogue: every sampler call returns this rank's rows of a stream
        that is the same for any number of ranks"""
        return self._rng


class UniformCatalog(RandomCatalog):
    """Poisson(nbar V) points uniform in the box, velocities uniform in 0.01 BoxSize"""

    def __repr__(self):
        return "UniformCatalog(size=%d, seed=%s)" % (self.size, self.attrs['seed'])

    @CurrentMPIComm.enable
    def __init__(self, nbar, BoxSize, seed=None, dtype='f8', comm=None):
        self.comm = comm
        box = numpy.empty(3, dtype='f8')
        box[:] = BoxSize
        self.attrs['BoxSize'] = box
        count = numpy.random.RandomState(seed).poisson(nbar * box.prod())
        if count == 0:
            raise ValueError("no uniform particles generated, try increasing `nbar` parameter")
        RandomCatalog.__init__(self, count, seed=seed, comm=comm)
        # the draw order is part of the stream: positions first, then velocities
        self._columns3 = {}
        for name, scale in (('Position', box), ('Velocity', 0.01 * box)):
            self._columns3[name] = (self.rng.uniform(itemshape=(3,)) * scale).astype(dtype)

    @column
    def Position(self):
        return self.make_column(self._columns3['Position'])

    @column
    def Velocity(self):
        return self.make_column(self._columns3['Velocity'])
