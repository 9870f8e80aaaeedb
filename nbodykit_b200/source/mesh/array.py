"""ArrayMesh -- a MeshSource from an in-memory array (API of nbodykit/source/mesh/array.py)."""
import numpy
import torch

from ... import CurrentMPIComm
from ...base.mesh import MeshSource
from ...pmesh.pm import RealField


class ArrayMesh(MeshSource):
    """
    array : ndarray or torch tensor, shape Nmesh (real) -- or the Hermitian-compressed complex array c of
            shape (Nx, Ny, Nz/2+1), interpreted as the field irfftn(c) * prod(Nmesh)  (array.py:36-37)
    BoxSize : float or 3-vector
    root : rank that holds the full array (other ranks may pass None)
    """

    def __repr__(self):
        return "ArrayMesh()"

    @CurrentMPIComm.enable
    def __init__(self, array, BoxSize, comm=None, root=0, **kwargs):
        if comm.rank == root:
            if isinstance(array, torch.Tensor):
                array = array.detach().cpu().numpy()
            array = numpy.array(array)
            if array.dtype.kind == 'c':
                array = numpy.fft.irfftn(array)
                array[...] *= numpy.prod(array.shape)
            shape, dtype = array.shape, array.dtype
        else:
            array, dtype, shape = [None] * 3
        dtype = comm.bcast(dtype, root=root)
        shape = comm.bcast(shape, root=root)
        assert len(shape) == 3, "only 3-D meshes are supported"
        MeshSource.__init__(self, comm, shape, BoxSize, numpy.empty(0, dtype).real.dtype)
        self.attrs.update(kwargs)
        self.field = RealField(self.pm)
        pm = self.pm
        if comm.size == 1:
            local = array
        else:
            parts = None
            if comm.rank == root:
                parts = [numpy.ascontiguousarray(array[r * pm.x_n:(r + 1) * pm.x_n]) for r in range(comm.size)]
            local = comm.bcast(parts, root=root)[comm.rank]
        self.field[...] = numpy.ascontiguousarray(local).astype(pm.dtype)

    def to_real_field(self):
        return self.field.copy()
