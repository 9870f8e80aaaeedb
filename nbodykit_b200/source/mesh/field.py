"""FieldMesh -- a MeshSource wrapping an in-memory (device) Field (API of nbodykit/source/mesh/field.py)."""
from ...base.mesh import MeshSource
from ...pmesh.pm import ComplexField, RealField


class FieldMesh(MeshSource):
    """the wrapped field is never modified: every conversion hands out a copy"""

    def __repr__(self):
        return "FieldMesh()"

    def __init__(self, field):
        MeshSource.__init__(self, field.pm.comm, field.Nmesh, field.BoxSize, field.pm.dtype)
        self.pm = field.pm      # share the decomposition of the wrapped field
        self.field = field

    def to_complex_field(self):
        if isinstance(self.field, ComplexField):
            return self.field.copy()
        return NotImplemented

    def to_real_field(self):
        if isinstance(self.field, RealField):
            return self.field.copy()
        return NotImplemented
