"""
FieldMesh -- a MeshSource around a device field that already exists (API of nbodykit/source/mesh/field.py:7-40).

The source shares the field's ParticleMesh (its slab decomposition and dtype) and answers only for the representation
the field is in; MeshSource.compute converts (one FFT) when the other one is asked for.  The wrapped field is never
modified: what leaves the source is a device-side copy.
"""
from ...base.mesh import MeshSource
from ...pmesh.pm import BaseComplexField, RealField


class FieldMesh(MeshSource):
    def __repr__(self):
        return "FieldMesh()"

    def __init__(self, field):
        pm = field.pm
        MeshSource.__init__(self, pm.comm, pm.Nmesh, pm.BoxSize, pm.dtype)
        self.pm = pm
        self.field = field

    def _copy_if(self, kind):
        return self.field.copy() if isinstance(self.field, kind) else NotImplemented

    def to_real_field(self):
        return self._copy_if(RealField)

    def to_complex_field(self):
        return self._copy_if(BaseComplexField)
