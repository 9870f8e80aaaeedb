from .catalog import CatalogMesh
from .field import FieldMesh
from .array import ArrayMesh
from .species import MultipleSpeciesCatalogMesh
from .linear import LinearMesh
from .file import FileMesh

BigFileMesh = FileMesh      # the role the reference's BigFileMesh plays (source/mesh/bigfile.py)

__all__ = ['CatalogMesh', 'FieldMesh', 'ArrayMesh', 'MultipleSpeciesCatalogMesh', 'LinearMesh', 'FileMesh', 'BigFileMesh']
