"""
LinearMesh -- a Gaussian random field with a given linear power spectrum, generated directly in Fourier space on the
device (API of nbodykit/source/mesh/linear.py:7-98; the field itself comes from mockmaker.gaussian_complex_fields).

The complex field is normalised to 1 + delta: its k = 0 mode is set to 1 (linear.py:84-96).  On several GPUs every rank
generates the same field (same seed) and keeps its own slab of the decomposition, so the realisation does not depend on
the number of ranks.
"""
import numbers

import numpy
import torch

from ... import CurrentMPIComm
from ...base.mesh import MeshSource
from ...comm import SelfComm
from ...pmesh.pm import ComplexField, ParticleMesh
from ... import mockmaker


class LinearMesh(MeshSource):
    def __repr__(self):
        return "LinearMesh(seed=%(seed)d)" % self.attrs

    @CurrentMPIComm.enable
    def __init__(self, Plin, BoxSize, Nmesh, seed=None, unitary_amplitude=False, inverted_phase=False,
                 remove_variance=None, comm=None):
        self.Plin = Plin
        if hasattr(Plin, 'attrs'):
            self.attrs.update(Plin.attrs)
        if seed is None:
            if comm.rank == 0:
                seed = numpy.random.randint(0, 4294967295)
            seed = comm.bcast(seed)
        if not isinstance(seed, numbers.Integral):
            raise ValueError("the seed used to generate the linear field must be an integer")
        self.attrs['seed'] = int(seed)
        if remove_variance is not None:
            unitary_amplitude = remove_variance
        self.attrs['unitary_amplitude'] = unitary_amplitude
        self.attrs['inverted_phase'] = inverted_phase
        MeshSource.__init__(self, BoxSize=BoxSize, Nmesh=Nmesh, dtype='f4', comm=comm)

    def to_complex_field(self):
        """the linear density field in Fourier space, normalised to 1 + delta"""
        pm = self.pm
        solo = pm if pm.comm.size == 1 else ParticleMesh(BoxSize=pm.BoxSize, Nmesh=pm.Nmesh, dtype='f4', comm=SelfComm())
        delta_k, _ = mockmaker.gaussian_complex_fields(solo, self.Plin, self.attrs['seed'],
                                                       unitary_amplitude=self.attrs['unitary_amplitude'],
                                                       inverted_phase=self.attrs['inverted_phase'],
                                                       compute_displacement=False)
        delta_k.value[0, 0, 0] = 1.0            # mean of the real-space field is unity
        if pm.comm.size > 1:
            # this rank's y-slab of the transposed layout [y_n][Nx][Nzc]
            out = ComplexField(pm)
            out.value.copy_(delta_k.value[:, pm.y_start:pm.y_start + pm.y_n, :].permute(1, 0, 2))
            delta_k = out
        delta_k.attrs = {}
        delta_k.attrs.update(self.attrs)
        return delta_k
